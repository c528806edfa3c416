"""Binary Sparse Coding with Expectation Truncation on the B200 engine.

Same constructor, methods, argument meaning and return dicts as
prosper/em/camodels/bsc_et.py (BSC_ET); each method cites the lines it replaces:
  select_Hprimes :98-115   E_step :119-192   M_step :195-438
"""
from math import pi as PI

import numpy as np
from scipy.special import comb

from ._gaussian_linear import GaussianLinearET
from ... import _lib
from ...utils.datalog import dlog


class BSC_ET(GaussianLinearET):
    model_kind = _lib.MODEL_BSC

    def __init__(self, D, H, Hprime, gamma, to_learn=['W', 'pi', 'sigma'], comm=None):
        GaussianLinearET.__init__(self, D, H, Hprime, gamma, to_learn, comm)

    def generate_from_hidden(self, model_params, my_hdata):
        """bsc_et.py:67-95: y = s.W^T + N(0, sigma) (host-side data generation)."""
        W = model_params['W'].T
        s = np.asarray(my_hdata['s'])
        y = s.astype(np.float64) @ W
        y += np.random.normal(scale=model_params['sigma'], size=y.shape)
        return {'y': y, 's': s}

    def _pack_params(self, model_params):
        if 'mu' not in model_params:          # bsc_et.py:145-149 inserts mu into the caller's dict
            model_params['mu'] = np.zeros(self.D)
        return self.engine.params(model_params['W'], model_params['pi'], model_params['sigma'],
                                  model_params.get('mu'))

    def _AB(self, pies):
        """A_pi_gamma, B_pi_gamma (bsc_et.py:239-243)."""
        H = self.H
        A = 0.
        B = 0.
        for g in range(self.gamma + 1):
            a = comb(H, g) * (pies ** g) * ((1 - pies) ** (H - g))
            A += a
            B += g * a
        return A, B

    def _truncation_mass(self, model_params):
        return self._AB(model_params['pi'])[0]

    def _log_before_L(self, N_use, A):
        dlog.append('N', N_use)                                          # bsc_et.py:261

    def _likelihood_const(self, model_params, A):
        pies, sigma = model_params['pi'], model_params['sigma']         # bsc_et.py:264
        return self.H * np.log(1 - pies) - 0.5 * self.D * np.log(2 * PI * sigma ** 2) - np.log(A)

    def _update_prior(self, model_params, counts, N_use, A):
        pies = model_params['pi']
        A, B = self._AB(pies)
        E = pies * self.H * A / B                                        # bsc_et.py:244
        return E * counts[0] / self.H / N_use                            # bsc_et.py:387

    def _result(self, model_params, W_new, pi_new, sigma_new):
        mu_new = model_params['mu']
        if 'mu' in self.to_learn:                                        # bsc_et.py:422-430
            # mu_new = sum_kept(y)/N - W_new . sum_kept(<s>)/N.  The reference divides by the LOCAL kept count
            # (`my_N`, :428), which is only meaningful on one rank; here N = N_use over all ranks (identical on one).
            stats, N_use, use_cut = self._mstep_ctx
            eng, lay = self.engine, self.engine.layout
            dsum = eng.data_sum(use_cut)                                 # of y - mu: the shard is stored shifted
            self.comm.allreduce_tensor_(dsum)
            mus = stats[lay.off_Wp + self.D * lay.ld_Wp:lay.off_Wp + self.D * lay.ld_Wp + self.H]   # sum_n <s> (ones row)
            data_sum = dsum.cpu().numpy() + N_use * np.asarray(model_params['mu'], dtype=np.float64)
            Wn = W_new.cpu().numpy() if hasattr(W_new, 'cpu') else W_new
            mu_new = data_sum / N_use - np.inner(Wn / N_use, mus.cpu().numpy())
        return {'W': W_new, 'pi': pi_new, 'sigma': sigma_new, 'mu': mu_new}
