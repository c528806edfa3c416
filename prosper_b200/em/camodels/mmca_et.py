"""Max-Magnitude Causes Analysis with Expectation Truncation on the B200 engine.

Mirrors prosper/em/camodels/mmca_et.py (MMCA_ET): check_params :48-60, generate_from_hidden :63-91,
select_Hprimes :95-124, E_step :127-202, M_step :205-426.
"""
import numpy as np

from ._max_causes import MaxCausesET
from . import CAModel
from ... import _lib


class MMCA_ET(MaxCausesET):
    model_kind = _lib.MODEL_MMCA

    def __init__(self, D, H, Hprime, gamma, to_learn=['W', 'pi', 'sigma'], comm=None):
        CAModel.__init__(self, D, H, Hprime, gamma, to_learn, comm)
        self.rho_T_bound = 1.20        # for rho: never use a T smaller than this
        self.rho_lbound = 1            # for rho: never use a rho smaller than this
        self.rho_ubound = 35           # for rho: never use a rho larger than this
        self.tol = 1e-4                # for W: ensure |W| >= tol
        tol = self.tol
        self.noise_policy = {
            'W': (-np.inf, +np.inf, False),
            'pi': (tol, 1 - tol, False),
            'sigma': (tol, +np.inf, False),
        }

    def check_params(self, model_params):
        """mmca_et.py:48-60: |W| >= tol, IN PLACE on the caller's array."""
        tol = self.tol
        W = model_params['W']
        W[np.logical_and(W >= 0., W < +tol)] = +tol
        W[np.logical_and(W <= 0., W > -tol)] = -tol
        return model_params

    def generate_from_hidden(self, model_params, my_hdata):
        """mmca_et.py:63-91: per feature the active cause of largest magnitude."""
        W, sigma = model_params['W'].T, model_params['sigma']
        s = np.asarray(my_hdata['s'])
        my_N = s.shape[0]
        y = np.zeros((my_N, self.D))
        for n in range(my_N):
            t0 = s[n, :, None] * W
            idx = np.argmax(np.abs(t0), axis=0)
            y[n] = t0[idx, np.arange(self.D)]
        y += np.random.normal(scale=sigma, size=(my_N, self.D))
        return {'y': y, 's': s}
