"""Shared host logic of the max-superposition ET models (MCA, MMCA).

Device pipeline: W^rho / log|W| tables -> score GEMM (singleton block) -> preselection ->
posterior kernel with a D-loop per state (`csrc/mca_kernel.cu`) -> statistics GEMM for the
singleton part + atomic scatter of the multi-cause `Aid` terms -> ONE all-reduce -> element-wise
W update.  Host arithmetic follows mca_et.py:182-377 / mmca_et.py:205-426.
"""
from math import pi as PI

import numpy as np
from scipy.special import comb

from . import CAModel
from ... import _lib
from ...utils.datalog import dlog
from ...utils import tracing


class MaxCausesET(CAModel):

    def _pack_params(self, model_params):
        return self.engine.params(model_params['W'], model_params['pi'], model_params['sigma'])

    def _latent_law(self, model_params):
        pi = float(model_params['pi'])                       # magnitude-max combination (mca_et.py:83-85)
        return np.array([0., 1.]), np.array([1. - pi, pi]), 1

    def _AB(self, pies):
        """mca_et.py:241-246 / mmca_et.py:269-274."""
        A = 0.
        B = 0.
        for gp in range(self.gamma + 1):
            a = comb(self.H, gp) * pies ** gp * (1. - pies) ** (self.H - gp)
            A += a
            B += gp * a
        return A, B

    @tracing.traced
    def select_Hprimes(self, model_params, data):
        self._bind(data)
        data['candidates'] = self.engine.select(self._pack_params(model_params))
        return data

    @tracing.traced
    def E_step(self, anneal, model_params, my_data):
        """-> {'logpj'}: NOT annealed (beta is applied in the M-step, mca_et.py:237-238)."""
        eng = self.engine
        self._bind(my_data)
        eng.set_candidates(my_data['candidates'])
        logpj = eng.e_step(eng.anneal(anneal), self._pack_params(model_params))
        assert np.isfinite(logpj).all()                                   # mca_et.py:177
        return {'logpj': logpj}

    @tracing.traced
    def M_step(self, anneal, model_params, my_suff_stat, my_data):
        eng = self.engine
        self._bind(my_data)
        eng.set_candidates(my_data['candidates'])
        logpj = np.ascontiguousarray(my_suff_stat['logpj'], dtype=np.float64)
        return self._m_step(anneal, model_params, logpj, fused=False)

    def _fused_step(self, anneal, model_params, my_data):
        self._bind(my_data)
        return self._m_step(anneal, model_params, None, fused=True)

    def _m_step(self, anneal, model_params, logpj, fused):
        comm, eng = self.comm, self.engine
        H, D = self.H, self.D
        p = self._pack_params(model_params)
        a = eng.anneal(anneal)
        pies, sigma = model_params['pi'], model_params['sigma']
        assert np.isfinite(np.log(pies / (1. - pies)))                    # mca_et.py:232
        N = comm.allreduce(eng.n)
        A, B = self._AB(pies)
        sel = _lib.PASS_SELECT if fused else 0
        N_use_target = int(N * (1 - (1 - A) * anneal['Ncut_factor'])) if anneal['Ncut_factor'] > 0.0 else 0
        if N_use_target > 0:            # mca_et.py:249-262 (annealed log-denominators); [-0] keeps every point
            lse = eng.log_denominators(a, p, logpj, sel)
            self._global_cut(lse, N_use_target)
            stats = eng.m_step_stats(a, p, logpj, _lib.PASS_REUSE_SCORES if fused else 0, use_cut=True)
        else:
            stats = eng.m_step_stats(a, p, logpj, sel)
        comm.allreduce_tensor_(stats)          # one collective: mca_et.py:208,262,340-341,357,366,371
        sc = eng.scalars(stats)
        N_use = int(round(sc[0]))
        dlog.append('N_use', N_use)                                       # mca_et.py:265
        if 'W' in self.to_learn:
            assert bool(np.isfinite(sc[:4]).all())                        # mca_et.py:336-337
            W_dev, _ = eng.solve(p, stats)
            W_new = W_dev.cpu().numpy()
        else:
            W_new = model_params['W']
        pi_new = A / B * pies * sc[3] / N_use if 'pi' in self.to_learn else pies          # mca_et.py:357
        sigma_new = np.sqrt(sc[2] / D / N_use) if 'sigma' in self.to_learn else sigma      # mca_et.py:366
        lAi = (H * np.log(1. - pi_new)) - ((D / 2) * np.log(2 * PI)) - (D * np.log(sigma_new))   # mca_et.py:372
        loglike_et = (lAi * N_use) + sc[1]                                                  # mca_et.py:375
        return {'W': W_new, 'pi': pi_new, 'sigma': sigma_new, 'Q': loglike_et}
