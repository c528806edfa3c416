"""Mixture of Poissons on the B200 engine; mirrors prosper/em/mixturemodels/MoP.py.

log p(y | h) = sum_d y_d log W_dh - W_dh (the -W term only without the normalisation constant A, MoP.py:204-207):
one GEMM Y . (log W)^T.  The reference accumulates these sums in np.float128; here they are FP64 tensor-core
sums (agreement ~1e-13 relative).  M-step: W_num = Y^T . posteriors (`pet_dgemm_mn`).
"""
import numpy as np

from . import MixtureModel

TINY = np.finfo(np.float64).tiny
EPS = np.finfo(np.float64).eps


class MoP(MixtureModel):
    def __init__(self, D, H, to_learn=['pies', 'W'], A=np.nan, comm=None):
        MixtureModel.__init__(self, D=D, H=H, to_learn=to_learn, comm=comm)
        if not np.isnan(A) and A <= D:
            A = 10 * D                                                      # MoP.py:24-25
        self.A = A

    def standard_init(self, my_data):
        return self.comm.bcast(MixtureModel.standard_init(self, my_data))    # MoP.py:29-40

    def generate_from_hidden(self, model_params, my_hdata):
        """MoP.py:65-87."""
        s = my_hdata['s']
        W = model_params['W']
        y = np.random.poisson(W[:, s].T).astype(np.float64)
        return {'y': y, 's': s}

    def check_params(self, model_params):
        assert np.isfinite(model_params['W']).all()
        assert np.isfinite(model_params['pies']).all()
        return model_params

    def normalize(self, my_y):
        """MoP.py:236-244 (host version, for callers that use it directly)."""
        my_y_sum = np.sum(my_y, 1) + EPS
        return ((self.A - self.D) / my_y_sum[:, None]) * my_y + 1

    # -- device paths ---------------------------------------------------------------------------------------
    def _data(self, my_y):
        """Device copy of the (normalised, if A is set) data."""
        Y, cache = self._bind(my_y)
        if np.isnan(self.A):
            return Y
        if 'Yn' not in cache:
            cache['Yn'] = self.ops.empty(my_y.shape[0], self.D)
            self.ops.rowop(3, my_y.shape[0], self.D, Y, None, 0, self.A - self.D, cache['Yn'])
        return cache['Yn']

    def _e_step_device(self, anneal, model_params, my_data):
        return self._posterior_device(model_params, my_data['y'], 1. / anneal['T'], _normalized=True)

    def _posterior_device(self, model_params, my_y, beta, _normalized=False):
        """MoP.py:165-217.  Called through E_step the data are normalised first (:97-98); the public
        posterior() takes them as given, like upstream."""
        ops, D, H = self.ops, self.D, self.H
        n = my_y.shape[0]
        Y = self._data(my_y) if _normalized else self._bind(my_y)[0]
        W = np.asarray(model_params['W'], dtype=np.float64)
        T1 = ops.empty(n, H)
        ops.gemm_kk(n, H, D, Y, ops.padded(np.log(W).T), T1)
        k = np.log(np.asarray(model_params['pies'], dtype=np.float64))
        if np.isnan(self.A):
            k = k - W.sum(axis=0)
        return ops.posterior(n, H, T1, None, 1.0, 0.0, k, beta)

    def _m_step_device(self, anneal, model_params, post, my_data):
        """MoP.py:102-157."""
        ops, D, H = self.ops, self.D, self.H
        my_y = my_data['y']
        n = my_y.shape[0]
        Y = self._data(my_y)
        Wnum = ops.empty(D, H)
        ops.gemm_mn(D, H, n, Y, post, Wnum)
        sum_post, Wnum = self._allreduce([ops.colsum(n, H, post), Wnum[:D]])
        sum_post = sum_post.cpu().numpy()
        if 'W' in self.to_learn:
            W_num = Wnum[:, :H].cpu().numpy()
            denom = sum_post if np.isnan(self.A) else np.sum(W_num, 0) / self.A + EPS
            model_params['W'] = (W_num / denom[None, :]) + EPS
        if 'pies' in self.to_learn:
            sp = sum_post + TINY
            model_params['pies'] = sp / np.sum(sp)
        return model_params
