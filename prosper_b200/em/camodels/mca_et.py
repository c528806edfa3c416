"""Maximal Causes Analysis with Expectation Truncation on the B200 engine.

Mirrors prosper/em/camodels/mca_et.py (MCA_ET): check_params :44-55, generate_data :58-86,
select_Hprimes :88-111, E_step :114-179, M_step :182-377.
"""
import numpy as np

from ._max_causes import MaxCausesET
from . import CAModel
from ... import _lib


class MCA_ET(MaxCausesET):
    model_kind = _lib.MODEL_MCA

    def __init__(self, D, H, Hprime, gamma, to_learn=['W', 'pi', 'sigma'], comm=None):
        CAModel.__init__(self, D, H, Hprime, gamma, to_learn, comm)
        self.rho_temp_bound = 1.05     # for rho: never use a T smaller than this
        self.W_tol = 1e-4              # for W: ensure W[W<W_tol] = W_tol
        W_tol = self.W_tol
        self.noise_policy = {
            'W': (W_tol, +np.inf, True),
            'pi': (W_tol, 1 - W_tol, False),
            'sigma': (W_tol, +np.inf, False),
        }

    def check_params(self, model_params):
        """mca_et.py:44-55: W >= W_tol (a NEW array, the caller's W is not modified)."""
        model_params = CAModel.check_params(self, model_params)
        model_params['W'] = np.maximum(model_params['W'], self.W_tol)
        return model_params

    def generate_data(self, model_params, my_N):
        """mca_et.py:58-86: max-rule superposition, not obeying gamma."""
        W, pies, sigma = model_params['W'].T, model_params['pi'], model_params['sigma']
        y = np.zeros((my_N, self.D))
        s = np.zeros((my_N, self.H), dtype=bool)
        for n in range(my_N):                      # one RNG call per datapoint, like the reference
            s[n] = np.random.random(self.H) < pies
            if s[n].any():
                y[n] = np.maximum(0., W[s[n]].max(axis=0))
        y += np.random.normal(scale=sigma, size=(my_N, self.D))
        return {'y': y, 's': s}

    def calculate_respons(self, anneal, model_params, data):
        """mca_et.py:380-387."""
        return self._responsibilities(anneal, model_params, data)
