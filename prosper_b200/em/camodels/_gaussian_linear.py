"""Shared host logic of the Gaussian-linear ET models (BSC, TSC, DSC).

All three evaluate ||y - sum_h s_h W_h||^2 over a truncated state space, so they share the
device pipeline (score GEMM -> posterior kernel -> statistics GEMM -> all-reduce -> solve);
what differs per model is O(1) host arithmetic: the truncation mass A, the likelihood
constant and the prior update.  Subclasses provide `_truncation_mass`, `_likelihood_const`,
`_update_prior` and `_pack_params`.
"""
import numpy as np
import torch

from . import CAModel
from ... import _lib
from ...utils.datalog import dlog
from ...utils import tracing


class GaussianLinearET(CAModel):
    # singular-value cutoff of the reference's solver: lstsq(rcond=-1) -> LAPACK machine epsilon 2^-53 on NumPy 2.x
    # (bsc_et.py:377-380, dsc_et.py:732-735); TSC overrides with pinv's default 1e-15 (tsc_et.py:493)
    _solve_rcond = 1.1102230246251565e-16

    # -- hooks ----------------------------------------------------------------------------------
    def _pack_params(self, model_params):
        return self.engine.params(model_params['W'], model_params['pi'], model_params['sigma'])

    def _truncation_mass(self, model_params):
        raise NotImplementedError

    def _likelihood_const(self, model_params, A):
        raise NotImplementedError

    def _update_prior(self, model_params, counts, N_use, A):
        raise NotImplementedError

    # -- the three operators ----------------------------------------------------------------------
    @tracing.traced
    def select_Hprimes(self, model_params, data):
        self._bind(data)
        data['candidates'] = self.engine.select(self._pack_params(model_params))
        return data

    @tracing.traced
    def E_step(self, anneal, model_params, my_data):
        eng = self.engine
        self._bind(my_data)
        eng.set_candidates(my_data['candidates'])
        return {'logpj': eng.e_step(eng.anneal(anneal), self._pack_params(model_params))}

    @tracing.traced
    def M_step(self, anneal, model_params, my_suff_stat, my_data):
        eng = self.engine
        self._bind(my_data)
        eng.set_candidates(my_data['candidates'])
        logpj = np.ascontiguousarray(my_suff_stat['logpj'], dtype=np.float64)
        return self._m_step(anneal, model_params, logpj, fused=False)

    def _fused_step(self, anneal, model_params, my_data):
        """select + E + M in one sweep over the shard; logpj (n x C) is never written to memory."""
        self._bind(my_data)
        return self._m_step(anneal, model_params, None, fused=True)

    def _m_step(self, anneal, model_params, logpj, fused):
        comm, eng = self.comm, self.engine
        p = self._pack_params(model_params)
        a = eng.anneal(anneal)
        A = self._truncation_mass(model_params)
        sel = _lib.PASS_SELECT if fused else 0
        self._check_data_noise(anneal)
        N_use_target = 0
        if anneal['Ncut_factor'] > 0.0:
            # N_use = int(N * (1 - (1 - A) * Ncut)); cut = allsort(denoms)[-N_use]   (bsc_et.py:250-252)
            N = comm.allreduce(eng.n)            # without a cut N = N_use comes back with the packed statistics
            N_use_target = int(N * (1 - (1 - A) * anneal['Ncut_factor']))
        if N_use_target > 0:                     # allsort(...)[-0] is the smallest denominator: N_use <= 0 keeps every point
            # fused: the same parameters come back for the statistics right after the cut -- the engine may evaluate the
            # posterior once and park the per-datapoint statistics (PET_PASS_DEFER_STATS)
            lse = eng.log_denominators(a, p, logpj, sel | (_lib.PASS_DEFER_STATS if fused else 0))
            self._global_cut(lse, N_use_target)
            stats = eng.m_step_stats(a, p, logpj, _lib.PASS_REUSE_SCORES if fused else 0, use_cut=True)
        else:
            stats = eng.m_step_stats(a, p, logpj, sel)
        comm.allreduce_tensor_(stats)     # ONE collective replaces bsc_et.py:258,266,373-374,387,417
        # the solve only needs the reduced statistics: enqueue it before the host waits for the scalars
        W_dev = None
        if 'W' in self.to_learn:
            W_dev, self.last_dropped_pivots = eng.solve(p, stats)
        sc = eng.scalars(stats)
        N_use = int(round(sc[0]))
        self._log_before_L(N_use, A)
        L = self._likelihood_const(model_params, A) + sc[1] / N_use
        dlog.append('L', L)
        if W_dev is not None:
            if self.last_dropped_pivots > 0:      # singular Wq: reproduce lstsq/pinv's minimum-norm answer
                W_dev = eng.solve_rank_deficient(stats, self._solve_rcond, self.model_kind == _lib.MODEL_BSC)
            # device-resident EM loop: parameters given as CUDA tensors come back as CUDA tensors; dlog
            # copies them to the host only if a handler subscribed to 'W' (datalog.py:215-232 `ignored()` pattern)
            W_new = W_dev if isinstance(model_params['W'], torch.Tensor) else W_dev.cpu().numpy()
        else:
            W_new = model_params['W']
        n_cnt = max(1, len(getattr(self, 'states', [0, 1])) - 1)
        pi_new = self._update_prior(model_params, sc[3:3 + n_cnt], N_use, A) if 'pi' in self.to_learn else model_params['pi']
        sigma_new = np.sqrt(sc[2] / self.D / N_use) if 'sigma' in self.to_learn else model_params['sigma']
        dlog.append('N_use', N_use)
        self._mstep_ctx = (stats, N_use, anneal['Ncut_factor'] > 0.0)
        return self._result(model_params, W_new, pi_new, sigma_new)

    def _log_before_L(self, N_use, A):
        pass

    @staticmethod
    def _check_data_noise(anneal):
        """bsc_et.py:228-230 adds my_data['data_noise'] to the M-step's copy of y when anneal['data_noise'] > 0.  The
        engine's shard is shared by the E- and the M-step, so that option is refused instead of silently ignored."""
        try:
            dn = anneal['data_noise']
        except (KeyError, IndexError):
            return
        if dn is not None and dn > 0:
            raise NotImplementedError("anneal['data_noise'] > 0 (bsc_et.py:228-230) is not supported on the device path")

    def _result(self, model_params, W_new, pi_new, sigma_new):
        return {'W': W_new, 'pi': pi_new, 'sigma': sigma_new, 'Q': 0.}
