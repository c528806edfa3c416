"""Discrete Sparse Coding with Expectation Truncation on the B200 engine.

Mirrors prosper/em/camodels/dsc_et.py (DSC_ET): select_Hprimes :347-410, E_step :492-585,
M_step :587-774, get_scaling_factors :798-823, truncation with strict '>' :825-843,
get_likelihood :845-870, standard_init :872-925.
"""
import itertools

import numpy as np
from scipy.special import gammaln

from ._gaussian_linear import GaussianLinearET
from . import CAModel
from ... import _lib
from ...utils import parallel
from ...utils.datalog import dlog


def get_states(states, Hprime, gamma):
    """dsc_et.py:56-63: product states with 2 <= nnz <= gamma, itertools.product order."""
    states = np.asarray(states)
    K = len(states)
    idx = np.indices((K,) * Hprime).reshape(Hprime, -1).T
    s = states[idx]
    nnz = (s != 0).sum(axis=1)
    return s[(nnz <= gamma) & (nnz > 1)]


class DSC_ET(GaussianLinearET):
    model_kind = _lib.MODEL_DSC

    def __init__(self, D, H, Hprime, gamma, states=np.array([-1., 0., 1.]), to_learn=['W', 'pi', 'sigma'], comm=None):
        CAModel.__init__(self, D, H, Hprime, gamma, to_learn, comm)
        if not type(states) == np.ndarray:
            raise TypeError("DSC: states must be of type numpy.ndarray")        # dsc_et.py:139
        if Hprime > H:
            raise Exception("Hprime must be less or equal to H")
        if gamma > Hprime:
            raise Exception("gamma must be less or equal to Hprime")
        self.states = states
        self.K = states.shape[0]
        self._K_0 = int(np.argwhere(states == 0.)[0, 0])
        self._noise_policy = dict(self.noise_policy)
        self.single_state_matrix = np.concatenate([np.eye(H) * states[i] for i in range(self.K) if i != self._K_0])
        self.state_matrix = get_states(states, Hprime, gamma)
        self.no_states = self.state_matrix.shape[0]
        self.state_abs = np.stack([(self.state_matrix == states[i]).sum(axis=1) for i in range(self.K)]).astype(np.float64)
        self.state_abs[self._K_0] = H - self.state_abs.sum(0) + self.state_abs[self._K_0]   # dsc_et.py:190-191

    def _make_engine(self):
        from . import Engine
        return Engine(self.model_kind, self.D, self.H, self.Hprime, self.gamma, states=self.states)

    def _latent_law(self, model_params):
        return np.asarray(self.states, dtype=np.float64), np.asarray(model_params['pi'], dtype=np.float64), 0

    # -- inference (dsc_et.py:927-1058) ---------------------------------------------------------
    def _regenerate_states(self):
        states, H = self.states, self.H
        self.state_matrix = get_states(states, self.Hprime, self.gamma)
        self.no_states = self.state_matrix.shape[0]
        self.state_abs = np.stack([(self.state_matrix == states[i]).sum(axis=1) for i in range(self.K)]).astype(np.float64)
        self.state_abs[self._K_0] = H - self.state_abs.sum(0) + self.state_abs[self._K_0]

    def _infer_fill(self, res, rows, idx, p, m, cand, logpj, topK, logprob):
        """dsc_et.py:996-1018: singleton columns carry their block's value; the marginal of a cause is the
        log-sum of its FIRST-block singleton and the multi-states holding the value 1 at it (reference quirk)."""
        H, nb = self.H, (self.K - 1) * self.H
        res['p'][rows] = p.cpu().numpy()
        res['m'][rows] = m.cpu().numpy()
        idx = idx.cpu().numpy().astype(np.int64)
        s = res['s']
        for k in range(topK):
            col = idx[:, k]
            single = (col >= 1) & (col < nb + 1)
            h = (col[single] - 1) % H
            s[rows[single], k, h] = self.single_state_matrix[col[single] - 1, h].astype(np.int8)
            multi = col >= nb + 1
            if multi.any():
                s[rows[multi][:, None], k, cand[multi]] = self.state_matrix[col[multi] - nb - 1].astype(np.int8)

    def check_params(self, model_params):
        """dsc_et.py:194-236."""
        assert np.isfinite(model_params['W']).all()
        assert np.isfinite(model_params['pi']).all()
        assert np.isfinite(model_params['sigma']).all()
        assert model_params['sigma'] >= 0.
        return model_params

    def generate_data(self, model_params, my_N, noise_on=True, gs=None, gp=None):
        """dsc_et.py:238-299: s_h ~ Categorical(pi) over `states` unless the latents are given (`gs`, optionally weighted
        by a posterior `gp` and summed over its first axis); `s` is stored as int8 like the reference's (a non-integer
        state value is truncated there too) and y is generated from that stored s; noise only if `noise_on`."""
        pi, W, sigma = model_params['pi'], model_params['W'].T, model_params['sigma']
        if gs is None:
            # one draw of my_N x H values consumes np.random exactly like the reference's my_N draws of H
            s = np.random.choice(self.states, size=(my_N, self.H), replace=True, p=pi).astype(np.int8)
        else:
            gs = np.asarray(gs)
            assert gs.shape[0] == my_N
            if gp is None:
                assert gs.ndim == 2
                s = gs.astype(np.int8)
            else:
                gp = np.asarray(gp)
                assert gp.shape[0] == my_N
                assert gp.shape[1] == gs.shape[1]
                s = (gs * gp).sum(1).astype(np.int8)
        y = np.dot(s, W).astype(np.float64)
        if noise_on:
            y += np.random.normal(scale=sigma, size=(my_N, self.D))
        return {'y': y, 's': s}

    def calculate_respons(self, anneal, model_params, data):
        """dsc_et.py:776-784."""
        return self._responsibilities(anneal, model_params, data)

    def free_energy(self, model_params, my_data):
        """dsc_et.py:786-790 (deprecated upstream)."""
        return 0.0

    def gain(self, old_parameters, new_parameters):
        """dsc_et.py:792-796 (deprecated upstream)."""
        return 0.0

    def _get_sorted_data(self, N, anneal, A_pi_gamma, all_denoms, candidates, logpj_all, my_y):
        """dsc_et.py:825-843: datapoint truncation of the compat path, strict `>` against the N_use-th largest
        denominator (the fused path does the same on the device: `pet_kth_largest` + the GLF_CUT_STRICT flag)."""
        if anneal['Ncut_factor'] > 0.0:
            N_use = int(N * (1 - (1 - A_pi_gamma) * anneal['Ncut_factor']))
            cut_denom = parallel.allsort(all_denoms, comm=self.comm)[-N_use]
            which = np.array(all_denoms > cut_denom)
            candidates, logpj_all, my_y = candidates[which], logpj_all[which], my_y[which]
            N_use = self.comm.allreduce(my_y.shape[0])
        else:
            N_use = N
        return N_use, my_y, candidates, logpj_all

    def get_likelihood(self, D, sigma, logpj_all, N):
        """dsc_et.py:845-870: -D/2 log(2 pi sigma^2) + sum_n logsumexp_c logpj[n, c] / N over all ranks."""
        import torch
        Fs = float(torch.logsumexp(torch.as_tensor(np.asarray(logpj_all, dtype=np.float64)), dim=1).sum())
        return -0.5 * D * np.log(2 * np.pi * sigma ** 2) + self.comm.allreduce(Fs) / N

    def noisify_params(self, model_params, anneal):
        """dsc_et.py:412-490: as the base class, except pi gets uniform noise and is renormalised."""
        comm = self.comm
        for param, policy in self._noise_policy.items():
            pvalue = model_params[param]
            scale = anneal[param + "_noise"]
            if scale == 0.0:
                continue
            if param == 'pi':
                new = pvalue
                if comm.rank == 0:
                    new = pvalue + np.random.rand(*pvalue.shape) * scale
                    new = new / new.sum()
                pvalue = comm.bcast(new)
            elif np.isscalar(pvalue):
                new = 0
                if comm.rank == 0:
                    new = pvalue + np.random.normal(scale=scale)
                    new = min(max(new, policy[0]), policy[1]) if new < policy[0] or new >= policy[1] else new
                    if policy[2]:
                        new = np.abs(new)
                pvalue = comm.bcast(new)
            else:
                new = pvalue
                if comm.rank == 0:
                    low, up, absify = policy
                    new = np.minimum(up, np.maximum(low, pvalue + np.random.normal(scale=scale, size=pvalue.shape)))
                    if absify:
                        new = np.abs(new)
                pvalue = comm.bcast(new)
            model_params[param] = pvalue
        return model_params

    def get_scaling_factors(self, pi):
        """Prior mass of the truncated state space (dsc_et.py:798-823)."""
        A = 0.0
        for gp in itertools.product(range(self.gamma + 1), repeat=self.K - 1):
            ngp = np.array(gp)
            if ngp.sum() > self.gamma:
                continue
            abs_array = np.insert(ngp, self._K_0, self.H - ngp.sum())
            cmb = np.exp(gammaln(abs_array.sum() + 1) - gammaln(abs_array + 1).sum())
            A += cmb * np.prod(pi ** abs_array)
        return A

    def _truncation_mass(self, model_params):
        A = self.get_scaling_factors(model_params['pi'])
        dlog.append("prior_mass", A)                                         # dsc_et.py:643
        return A

    def _likelihood_const(self, model_params, A):
        sigma = model_params['sigma']                                        # dsc_et.py:866 (no -log A)
        return -0.5 * self.D * np.log(2 * np.pi * sigma ** 2)

    def _update_prior(self, model_params, counts, N_use, A):
        my_pi = np.zeros(self.K)
        nz = [i for i in range(self.K) if i != self._K_0]
        my_pi[nz] = counts[:len(nz)]
        my_pi[self._K_0] = self.H * N_use - my_pi.sum()      # every state has H entries and posteriors sum to 1
        pi_new = my_pi / my_pi.sum()                                         # dsc_et.py:740-741
        eps = 1e-6                                                           # dsc_et.py:743-748
        if np.any(pi_new < eps):
            lo = pi_new < eps
            hi = ~lo
            pi_new[lo] += eps - pi_new[lo]
            pi_new[hi] -= (eps * lo.sum()) / hi.sum()
        if 'penalty' in self.__dict__:                                       # dsc_et.py:750-756
            if self.penalty > pi_new[self._K_0]:
                r = (1 - self.penalty) / (1 - pi_new[self._K_0])
                pi_new[pi_new != 0] = pi_new[pi_new != 0] * r
                pi_new[self._K_0] = self.penalty
                pi_new /= pi_new.sum()
        return pi_new

    def standard_init(self, data):
        """dsc_et.py:872-925: W/sigma as the base class, pi = sparsity 1-1/H on the zero state."""
        comm = self.comm
        base = CAModel.standard_init(self, data)
        sparsity = 1. - (1. / self.H)
        pi_init = np.random.rand(self.K - 1) if comm.rank == 0 else None
        pi_init = comm.bcast(pi_init)
        pi_init = (1 - sparsity) * pi_init / pi_init.sum()
        base['pi'] = np.insert(pi_init, self._K_0, sparsity)
        return base
