"""Ternary Sparse Coding with Expectation Truncation on the B200 engine.

Mirrors prosper/em/camodels/tsc_et.py (TSC_ET): select_Hprimes :142-212, E_step :277-356,
M_step :359-542.  The reference's constructor raises NameError upstream (`states` undefined,
:131); here the ternary values are fixed to [-1, 0, 1] as :125 intends.
"""
import numpy as np
import torch
from scipy.special import comb

from . import CAModel
from ._gaussian_linear import GaussianLinearET
from .. import Model
from ... import _lib


def generate_state_matrix(Hprime, gamma, H, states):
    """tsc_et.py:23-80 -> (single_state_matrix, state_matrix, no_states (= K**H', unfiltered), states_abs)."""
    states = np.asarray(states, dtype=np.float64)
    ssm = np.concatenate([np.eye(H, dtype=np.int8) * int(v) for v in states if v != 0])
    K = len(states)
    idx = np.indices((K,) * Hprime).reshape(Hprime, -1).T
    s = states[idx].astype(np.int8)
    states_abs = np.stack([(s == v).sum(axis=1) for v in states]).astype(np.float64)
    return ssm, s[np.abs(s).sum(axis=1) <= gamma], s.shape[0], states_abs


class TSC_ET(GaussianLinearET):
    model_kind = _lib.MODEL_TSC
    _solve_rcond = 1e-15          # np.linalg.pinv default (tsc_et.py:493)

    def __init__(self, D, H, Hprime, gamma, to_learn=['W', 'pi', 'sigma'], comm=None):
        Model.__init__(self, comm)
        self.to_learn = to_learn
        self.states = np.array([-1., 0., 1.])
        self.gamma, self.D, self.H, self.Hprime = gamma, D, H, Hprime
        assert Hprime <= H and gamma <= Hprime
        self.single_state_matrix, self.state_matrix, self.no_states, self.state_abs = \
            generate_state_matrix(Hprime, gamma, H, self.states)
        tol = 1e-5
        self.noise_policy = {'W': (-np.inf, +np.inf, False), 'pi': (tol, 1. - tol, False), 'sigma': (0., +np.inf, False)}
        self.cache_data = True
        self._engine = None
        self._bound = None

    def generate_data(self, model_params, my_N):
        """tsc_et.py:214-275: s in {-1,0,1} with p(-1)=p(+1)=pi/2."""
        pi, W, sigma = model_params['pi'], model_params['W'].T, model_params['sigma']
        p = np.random.random((my_N, self.H))
        s = np.where(p < pi / 2, -1, np.where(p < pi, 1, 0)).astype(np.int8)
        y = s.astype(np.float64) @ W + np.random.normal(scale=sigma, size=(my_N, self.D))
        return {'y': y, 's': s}

    def _latent_law(self, model_params):
        pi = float(model_params['pi'])                       # tsc_et.py:228-236
        return np.array([-1., 0., 1.]), np.array([pi / 2, 1. - pi, pi / 2]), 0

    # -- inference (tsc_et.py:546-680) ----------------------------------------------------------
    def _regenerate_states(self):
        self.single_state_matrix, self.state_matrix, self.no_states, self.state_abs = \
            generate_state_matrix(self.Hprime, self.gamma, self.H, self.states)

    def inference(self, anneal, model_params, test_data, topK=10, logprob=False, abs_marginal=True,
                  adaptive=True, Hprime_max=None, gamma_max=None):
        """tsc_et.py:546-547: the reference's positional order (`abs_marginal` sits before `adaptive`); adds res['am']."""
        return CAModel.inference(self, anneal, model_params, test_data, topK=topK, logprob=logprob, adaptive=adaptive,
                                 Hprime_max=Hprime_max, gamma_max=gamma_max, abs_marginal=abs_marginal)

    def _infer_res(self, my_N, topK):
        res = CAModel._infer_res(self, my_N, topK)
        res['am'] = np.zeros((my_N, self.H))
        return res

    def _infer_marginals(self):
        return False                 # TSC returns posterior means instead of per-cause log-marginals

    def _infer_kernel_logprob(self, logprob):
        return True                  # p is the NORMALISED probability here (tsc_et.py:632)

    def _infer_fill(self, res, rows, idx, p, m, cand, logpj, topK, logprob):
        import ctypes as C
        eng = self.engine
        n, S = logpj.shape
        res['p'][rows] = p.cpu().numpy() if logprob else np.exp(p.cpu().numpy())
        idx = idx.cpu().numpy().astype(np.int64)
        for k in range(topK):                                                   # :633-634
            res['s'][rows[:, None], k, cand] = self.state_matrix[idx[:, k]].astype(np.int8)
        # posterior means over the candidates: (P . SM) and (P . |SM|), one DMMA GEMM (tsc_et.py:636-638)
        ldp = (S + 1) // 2 * 2
        P = torch.zeros((n, ldp), dtype=torch.float64, device=logpj.device)
        P[:, :S] = torch.softmax(logpj, dim=1)
        Hp = self.Hprime
        B = torch.zeros((2 * Hp, ldp), dtype=torch.float64, device=logpj.device)
        sm = torch.as_tensor(np.ascontiguousarray(self.state_matrix.T, dtype=np.float64)).to(logpj.device)
        B[:Hp, :S] = sm
        B[Hp:, :S] = sm.abs()
        out = torch.empty((n, 2 * Hp), dtype=torch.float64, device=logpj.device)
        _lib.check(eng.lib.pet_dgemm_kk(n, 2 * Hp, S, C.c_void_p(P.data_ptr()), ldp, C.c_void_p(B.data_ptr()), ldp,
                                        C.c_void_p(out.data_ptr()), 2 * Hp, 1.0, 0.0, eng.stream()))
        out = out.cpu().numpy()
        res['m'][rows[:, None], cand] = out[:, :Hp]                            # duplicates: last write wins
        if self._infer_kwargs.get('abs_marginal', True):
            res['am'][rows[:, None], cand] = out[:, Hp:]

    def _infer_finish(self, res, logprob):
        if logprob:                                                             # :671-673
            with np.errstate(divide='ignore', invalid='ignore'):
                res['m'] = np.log(res['m'])
                res['am'] = np.log(res['am'])

    def _infer_map_activity(self, res):
        return (res['s'][:, 0, :].astype(bool) != 0).sum(-1)                   # :643

    def _AB(self, pi):
        """Trinomial truncation sums (tsc_et.py:422-431)."""
        H, gamma = self.H, self.gamma
        A = 0.0
        B = 0.0
        for g1 in range(gamma + 1):
            for g2 in range(gamma - g1 + 1):
                cmb = comb(g1, g1) * comb(g1 + g2, g2) * comb(H, H - g1 - g2)
                a = cmb * ((pi / 2) ** (g1 + g2)) * ((1 - pi) ** (H - g1 - g2))
                A += a
                B += (g1 + g2) * a
        return A, B

    def _truncation_mass(self, model_params):
        return self._AB(model_params['pi'])[0]

    def _likelihood_const(self, model_params, A):
        sigma = model_params['sigma']                                    # tsc_et.py:447 (no H*log(1-pi) term)
        return -0.5 * self.D * np.log(2 * np.pi * sigma ** 2) - np.log(A)

    def _update_prior(self, model_params, counts, N_use, A):
        pi = model_params['pi']
        A, B = self._AB(pi)
        E = pi * self.H * A / B
        return E * (counts[0] + counts[1]) / self.H / N_use              # tsc_et.py:479,501
