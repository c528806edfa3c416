"""Mixture of Gaussians on the B200 engine; mirrors prosper/em/mixturemodels/MoG.py.

log p(y | h) follows the reference literally: -(log det Sigma_h + (y - W_h)^T Sigma_h^-1 (y - W_h)) * beta, i.e.
without the factor 1/2 of a Gaussian density (MoG.py:255,259).  'diagonal': the quadratic form expands into two
GEMMs, (Y o Y) . (1/sigma^2)^T and Y . (W / sigma^2)^T; 'full': per component one GEMM (Y - W_h) . Sigma_h^-1 and a
row dot.  The (H, D, D) inverses / log-determinants are data independent and computed with torch.linalg on the
device.  M-step: W and the second moments are `pet_dgemm_mn` products with the posterior matrix.
"""
import numpy as np
import torch

from . import MixtureModel

TINY = np.finfo(np.float64).tiny


class MoG(MixtureModel):
    def __init__(self, D, H, to_learn=['pies', 'W', 'sigmas_sq'], sigmas_sq_type='full', comm=None):
        MixtureModel.__init__(self, D=D, H=H, to_learn=to_learn, comm=comm)
        self.sigmas_sq_type = sigmas_sq_type

    def standard_init(self, my_data):
        """MoG.py:23-57."""
        comm = self.comm
        H = self.H
        my_y = my_data['y']
        N, D = my_y.shape
        model_params = MixtureModel.standard_init(self, my_data)
        if 'sigmas_sq' in self.to_learn:
            if self.sigmas_sq_type == 'full':
                sigma = comm.bcast(np.cov(my_y.T) + (0.001 * np.eye(D)))
            elif self.sigmas_sq_type == 'diagonal':
                sigma = comm.bcast(np.var(my_y, axis=0) + 0.001)
            sigmas_sq = np.zeros(tuple([H]) + sigma.shape)
            for h in range(H):
                sigmas_sq[h] = sigma
            model_params['sigmas_sq'] = sigmas_sq
        return comm.bcast(model_params)

    def generate_from_hidden(self, model_params, my_hdata):
        """MoG.py:104-131 (vectorised; same draws per datapoint as the upstream loop)."""
        s = my_hdata['s']
        my_N = s.size
        W = model_params['W'].T
        y = np.zeros((my_N, self.D))
        for n in range(my_N):
            comp = s[n]
            sigma = model_params['sigmas_sq'][comp].diagonal() if self.sigmas_sq_type == 'full' else model_params['sigmas_sq'][comp]
            y[n] = W[comp] + np.sqrt(sigma) * np.random.randn(self.D)
        return {'y': y, 's': s}

    def check_params(self, model_params):
        assert np.isfinite(model_params['W']).all()
        assert np.isfinite(model_params['sigmas_sq']).all()
        assert np.isfinite(model_params['pies']).all()
        return model_params

    # -- device paths ---------------------------------------------------------------------------------------
    def _e_step_device(self, anneal, model_params, my_data):
        return self._posterior_device(model_params, my_data['y'], 1. / anneal['T'])       # MoG.py:133-141

    def _posterior_device(self, model_params, my_y, beta):
        """MoG.py:208-262 -> (logpj, posteriors) on the device."""
        ops, D, H = self.ops, self.D, self.H
        Y, cache = self._bind(my_y)
        n = my_y.shape[0]
        W = np.asarray(model_params['W'], dtype=np.float64)                 # (D, H)
        logpies = np.log(np.asarray(model_params['pies'], dtype=np.float64))
        if self.sigmas_sq_type == 'diagonal':
            sig = np.asarray(model_params['sigmas_sq'], dtype=np.float64)     # (H, D)
            iv = 1. / sig
            if 'Y2' not in cache:
                cache['Y2'] = ops.empty(n, D)
                ops.rowop(0, n, D, Y, None, 0, 0.0, cache['Y2'])
            T1, T2 = ops.empty(n, H), ops.empty(n, H)
            ops.gemm_kk(n, H, D, cache['Y2'], ops.padded(iv), T1)             # sum_d y^2 / sigma^2
            ops.gemm_kk(n, H, D, Y, ops.padded(W.T * iv), T2)                 # sum_d y W / sigma^2
            k = -(np.log(sig).sum(axis=1) + (W.T ** 2 * iv).sum(axis=1)) + logpies
            return ops.posterior(n, H, T1, T2, -1.0, 2.0, k, beta)
        sig = torch.as_tensor(np.ascontiguousarray(model_params['sigmas_sq'], dtype=np.float64)).to(ops.dev)   # (H, D, D)
        inv = torch.linalg.inv(sig)
        logdet = torch.linalg.slogdet(sig)[1].cpu().numpy()
        T1 = ops.empty(n, H)
        Yc, T = ops.empty(n, D), ops.empty(n, D)
        Wt = ops.padded(W.T)
        for h in range(H):
            A = torch.zeros((D, Yc.stride(0)), dtype=torch.float64, device=ops.dev)
            A[:, :D] = inv[h].T                                               # gemm_kk wants B[e][d] = A_h[d][e]
            ops.rowop(2, n, D, Y, Wt[h], 0, 0.0, Yc)                          # y - W_h
            ops.gemm_kk(n, D, D, Yc, A, T)                                    # (y - W_h) Sigma_h^-1
            ops.rowdot(n, D, T, Yc, T1[:, h], T1.stride(0))
        return ops.posterior(n, H, T1, None, -1.0, 0.0, -logdet + logpies, beta)

    def _m_step_device(self, anneal, model_params, post, my_data):
        """MoG.py:143-198 (mutates and returns model_params like upstream)."""
        ops, D, H = self.ops, self.D, self.H
        my_y = my_data['y']
        Y, cache = self._bind(my_y)
        n = my_y.shape[0]
        stats = [ops.colsum(n, H, post)]
        if 'W' in self.to_learn:
            Wnum = ops.empty(D, H)
            ops.gemm_mn(D, H, n, Y, post, Wnum)
            stats.append(Wnum[:D])
        if 'sigmas_sq' in self.to_learn:
            if self.sigmas_sq_type == 'diagonal':
                if 'Y2' not in cache:
                    cache['Y2'] = ops.empty(n, D)
                    ops.rowop(0, n, D, Y, None, 0, 0.0, cache['Y2'])
                S2 = ops.empty(D, H)
                ops.gemm_mn(D, H, n, cache['Y2'], post, S2)
                stats.append(S2[:D])
            else:
                C_all = torch.empty((H, D, _ld(D)), dtype=torch.float64, device=ops.dev)
                Yh = ops.empty(n, D)
                for h in range(H):
                    ops.rowop(1, n, D, Y, post[:, h], post.stride(0), 0.0, Yh)
                    ops.gemm_mn(D, D, n, Y, Yh, C_all[h])
                stats.append(C_all)
        stats = self._allreduce(stats)
        sum_post = stats[0].cpu().numpy() + TINY
        i = 1
        if 'W' in self.to_learn:
            model_params['W'] = stats[i][:, :H].cpu().numpy() * np.power(sum_post, -1)[None, :]
            i += 1
        if 'sigmas_sq' in self.to_learn:
            Wn = model_params['W']
            if self.sigmas_sq_type == 'diagonal':
                model_params['sigmas_sq'] = (stats[i][:, :H].cpu().numpy().T * np.power(sum_post, -1)[:, None]) - Wn.T ** 2
            else:
                sig = stats[i][:, :, :D].cpu().numpy() * np.power(sum_post, -1)[:, None, None]
                for h in range(H):
                    sig[h] -= np.outer(Wn[:, h], Wn[:, h])
                model_params['sigmas_sq'] = sig
        if 'pies' in self.to_learn:
            model_params['pies'] = sum_post / np.sum(sum_post)
        return model_params


def _ld(x):
    return (x + 1) // 2 * 2
