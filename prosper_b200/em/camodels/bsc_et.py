"""Binary Sparse Coding with Expectation Truncation on the B200 engine.

Same constructor, methods, argument meaning and return dicts as
prosper/em/camodels/bsc_et.py (BSC_ET); each method cites the lines it replaces.
"""
from math import pi as PI

import numpy as np
import torch
from scipy.special import comb

from . import CAModel
from ... import _lib
from ...utils.datalog import dlog


class BSC_ET(CAModel):
    model_kind = _lib.MODEL_BSC

    def __init__(self, D, H, Hprime, gamma, to_learn=['W', 'pi', 'sigma'], comm=None):
        CAModel.__init__(self, D, H, Hprime, gamma, to_learn, comm)

    def generate_from_hidden(self, model_params, my_hdata):
        """bsc_et.py:67-95: y = s.W^T + N(0, sigma) (host-side data generation)."""
        W = model_params['W'].T
        s = np.asarray(my_hdata['s'])
        y = s.astype(np.float64) @ W
        y += np.random.normal(scale=model_params['sigma'], size=y.shape)
        return {'y': y, 's': s}

    # -- helpers ----------------------------------------------------------------------------
    def _params(self, model_params):
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

    # -- the three operators ----------------------------------------------------------------
    def select_Hprimes(self, model_params, data):
        """bsc_et.py:98-115 -> data['candidates'] (n,H') int64, ascending by cosine score."""
        self._bind(data)
        data['candidates'] = self.engine.select(self._params(model_params))
        return data

    def E_step(self, anneal, model_params, my_data):
        """bsc_et.py:119-192 -> {'logpj': (n, 1+H+no_states)} (host array; compat path)."""
        eng = self.engine
        if self._bind(my_data) or 'candidates' in my_data:
            eng.set_candidates(my_data['candidates'])
        logpj = eng.e_step(eng.anneal(anneal), self._params(model_params))
        return {'logpj': logpj}

    def M_step(self, anneal, model_params, my_suff_stat, my_data):
        """bsc_et.py:195-438 with the caller's logpj (compat path)."""
        eng = self.engine
        if self._bind(my_data) or 'candidates' in my_data:
            eng.set_candidates(my_data['candidates'])
        logpj = np.ascontiguousarray(my_suff_stat['logpj'], dtype=np.float64)
        return self._m_step(anneal, model_params, logpj, fused=False)

    def _fused_step(self, anneal, model_params, my_data):
        """select + E + M in one sweep: logpj (n x C) is never written to memory."""
        self._bind(my_data)
        return self._m_step(anneal, model_params, None, fused=True)

    def _m_step(self, anneal, model_params, logpj, fused):
        comm, eng = self.comm, self.engine
        H, D = self.H, self.D
        p = self._params(model_params)
        a = eng.anneal(anneal)
        pies, sigma, mu = model_params['pi'], model_params['sigma'], model_params['mu']
        my_N = eng.n
        N = comm.allreduce(my_N)                                         # bsc_et.py:225
        A, B = self._AB(pies)
        E = pies * H * A / B                                             # :244

        sel = _lib.PASS_SELECT if fused else 0
        if anneal['Ncut_factor'] > 0.0:                                  # :247-258
            N_use_target = int(N * (1 - (1 - A) * anneal['Ncut_factor']))
            lse = eng.log_denominators(a, p, logpj, sel)
            self._global_cut(lse, N_use_target)
            stats = eng.m_step_stats(a, p, logpj, _lib.PASS_REUSE_SCORES if fused else 0, use_cut=True)
        else:
            stats = eng.m_step_stats(a, p, logpj, sel)
        if fused:
            self.last_candidates_on_device = True
        comm.allreduce_tensor_(stats)        # ONE collective: bsc_et.py:258,266,373-374,387,417
        sc = eng.scalars(stats)
        N_use = int(round(sc[0]))
        dlog.append('N', N_use)                                          # :261

        L = H * np.log(1 - pies) - 0.5 * D * np.log(2 * PI * sigma ** 2) - np.log(A)   # :264
        L += sc[1] / N_use                                               # :265-266
        dlog.append('L', L)

        if 'W' in self.to_learn:                                         # :369-382
            W_dev, self.last_dropped_pivots = eng.solve(p, stats)
            W_new = W_dev.cpu().numpy()
        else:
            W_new = model_params['W']
        pi_new = E * sc[3] / H / N_use if 'pi' in self.to_learn else pies            # :385-389
        sigma_new = np.sqrt(sc[2] / D / N_use) if 'sigma' in self.to_learn else sigma   # :417
        if 'mu' in self.to_learn:                                        # :422-430 (divides by the LOCAL my_N)
            lay = eng.layout
            mus = stats[lay.off_Wp + D * lay.ld_Wp: lay.off_Wp + D * lay.ld_Wp + H].cpu().numpy()
            raise NotImplementedError("learning 'mu' needs the data sum of the kept datapoints; not built yet")
        mu_new = mu
        dlog.append('N_use', N_use)                                      # :436
        return {'W': W_new, 'pi': pi_new, 'sigma': sigma_new, 'mu': mu_new}
