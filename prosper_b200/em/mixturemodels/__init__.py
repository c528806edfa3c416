"""Mixture models on the B200 engine.

Mirrors prosper/em/mixturemodels/__init__.py (MixtureModel :21-153).  The dense (n, H) posterior and the
sufficient statistics run on the device: the contractions are the engine's FP64 tensor-core GEMMs
(`pet_dgemm_kk`, `pet_dgemm_mn`), the posterior / element-wise pieces are `csrc/mixture.cu`.  Like the
CAModel mirror, the three operators keep the reference's signatures and return host NumPy arrays, and
`step` runs E and M back to back without the round trip.  There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from .. import Model
from ... import _lib
from ...utils import parallel
from ...utils.datalog import dlog
from ...utils import tracing


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _even(x):
    return (x + 1) // 2 * 2


class DeviceOps(object):
    """Thin helper over the C ABI for dense FP64 work on one device (all tensors row-major, even leading dims)."""

    def __init__(self, device=None):
        assert torch.cuda.is_available(), "prosper_b200 has no CPU path"
        self.lib = _lib.load()
        self.dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def padded(self, a):
        """Host (r, c) array -> device (r, even(c)) tensor, zero padded."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        t = torch.zeros((a.shape[0], _even(a.shape[1])), dtype=torch.float64, device=self.dev)
        t[:, :a.shape[1]] = torch.as_tensor(a)
        return t

    def empty(self, r, c):
        return torch.empty((max(r, 1), _even(c)), dtype=torch.float64, device=self.dev)

    def gemm_kk(self, M, N, K, A, B, out):
        """out[:M, :N] = A[:M, :K] . B[:N, :K]^T"""
        _lib.check(self.lib.pet_dgemm_kk(M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(out), out.stride(0), 1.0, 0.0, self.stream()))

    def gemm_mn(self, M, N, K, A, B, out):
        """out[:M, :N] = A[:K, :M]^T . B[:K, :N] (reduction over rows)"""
        splits = self.lib.pet_dgemm_mn(M, N, K, None, A.stride(0), None, B.stride(0), None, out.stride(0), 0, None, 0, self.stream())
        work = torch.empty(max(1, splits * M * out.stride(0)), dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.pet_dgemm_mn(M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(out), out.stride(0), 0, _p(work),
                                         work.numel(), self.stream()))

    def rowop(self, op, n, D, X, w, wstride, a, out):
        _lib.check(self.lib.pet_rowop(op, n, D, _p(X), X.stride(0), _p(w), wstride, float(a), _p(out), out.stride(0), self.stream()))

    def rowdot(self, n, D, A, B, out, stride):
        _lib.check(self.lib.pet_rowdot(n, D, _p(A), A.stride(0), _p(B), B.stride(0), _p(out), stride, self.stream()))

    def colsum(self, n, cols, M):
        out = torch.zeros(_even(cols), dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.pet_colsum(n, cols, _p(M), M.stride(0), _p(out), self.stream()))
        return out[:cols]

    def posterior(self, n, H, T1, T2, s1, s2, k, beta):
        """-> (logpj, posteriors) device tensors (n, even(H))"""
        lp, post = self.empty(n, H), self.empty(n, H)
        kd = torch.as_tensor(np.ascontiguousarray(k, dtype=np.float64)).to(self.dev)
        _lib.check(self.lib.pet_mix_posterior(n, H, _p(T1), _p(T2), T1.stride(0), float(s1), float(s2), _p(kd), float(beta),
                                              _p(lp), lp.stride(0), _p(post), post.stride(0), self.stream()))
        return lp, post


class MixtureModel(Model):
    """prosper/em/mixturemodels/__init__.py:21-153."""

    def __init__(self, D, H, to_learn=['W', 'pies'], comm=None):
        Model.__init__(self, comm)
        self.to_learn = to_learn
        self.D = D
        self.H = H
        self._ops = None
        self._bound = None

    @property
    def ops(self):
        if self._ops is None:
            self._ops = DeviceOps()
        return self._ops

    def invalidate_data(self):
        """Forget the device copy of the data (call after refilling a data buffer in place)."""
        self._bound = None

    def _bind(self, my_y):
        """Device copy of the data (n, even(D)), cached on the IDENTITY of the host array: the cache entry keeps a
        reference to the array it was made from, so another batch that happens to land at the same address with
        the same shape (e.g. successive `data[idx]` temporaries) is a different object and is uploaded again; a
        buffer refilled in place is caught by the content fingerprint of CAModel._data_key."""
        from ..camodels import CAModel
        key = CAModel._data_key(my_y)
        if self._bound is None or self._bound[0] != key or self._bound[3] is not my_y:
            host = my_y.cpu().numpy() if isinstance(my_y, torch.Tensor) else my_y
            self._bound = (key, self.ops.padded(host), {}, my_y)
        return self._bound[1], self._bound[2]

    def standard_init(self, data):
        """:33-63."""
        comm = self.comm
        H = self.H
        my_y = data['y']
        my_N, D = my_y.shape
        assert D == self.D
        W_mean = parallel.allmean(my_y, axis=0, comm=comm)
        sigma_sq = parallel.allmean((my_y - W_mean) ** 2, axis=0, comm=comm)
        sigma_init = np.sqrt(sigma_sq).sum() / D
        noise = sigma_init / 4.
        W_init = W_mean[:, None] + np.random.normal(scale=noise, size=[D, H])
        model_params = {'W': W_init}
        if 'pies' in self.to_learn:
            model_params['pies'] = np.ones(H) * 1. / H
        return model_params

    def check_params(self, model_params):
        raise NotImplementedError

    def generate_data(self, model_params, my_N):
        """:75-88 (component index drawn from `pies`, then generate_from_hidden)."""
        s = np.random.choice(self.H, size=my_N, p=np.asarray(model_params['pies']) / np.sum(model_params['pies']))
        return self.generate_from_hidden(model_params, {'s': s})

    def select_partial_data(self, anneal, data):
        """:90-113 (as upstream this indexes `data` itself, so it only works for partial in {0, 1} with a dict)."""
        partial = anneal['partial']
        if partial == 0 or partial == 1:
            return data
        my_N, D = data.shape
        my_pN = int(np.ceil(my_N * partial))
        if my_N == my_pN:
            return data
        sel = np.random.permutation(my_N)[:my_pN]
        return data[sel]

    @tracing.traced
    def step(self, anneal, model_params, data):
        """:115-138; E and M back to back on the device (the posterior never visits the host)."""
        model_params = self.noisify_params(model_params, anneal)
        model_params = self.check_params(model_params)
        pdata = self.select_partial_data(anneal, data)
        post_dev = self._e_step_device(anneal, model_params, pdata)
        new_model_params = self._m_step_device(anneal, model_params, post_dev[1], pdata)
        dlog.append_all(new_model_params)
        dlog.append_all(anneal.as_dict())
        return new_model_params

    @tracing.traced
    def E_step(self, anneal, model_params, my_data):
        lp, post = self._e_step_device(anneal, model_params, my_data)
        n = my_data['y'].shape[0]
        return {'posteriors_h': post[:n, :self.H].cpu().numpy(), 'logpj': lp[:n, :self.H].cpu().numpy()}

    @tracing.traced
    def M_step(self, anneal, model_params, suff_stats, my_data):
        post = self.ops.padded(suff_stats['posteriors_h'])
        return self._m_step_device(anneal, model_params, post, my_data)

    def posterior(self, model_params, my_y, beta=1.0):
        lp, post = self._posterior_device(model_params, my_y, beta)
        n = my_y.shape[0]
        return {'posteriors_h': post[:n, :self.H].cpu().numpy(), 'logpj': lp[:n, :self.H].cpu().numpy()}

    def inference(self, anneal, model_params, my_data, no_maps=10):
        """To be implemented (upstream stub, :140-143)."""

    # packed all-reduce of the M-step statistics
    def _allreduce(self, tensors):
        if self.comm.size == 1:
            return tensors
        flat = torch.cat([t.reshape(-1) for t in tensors])
        self.comm.allreduce_tensor_(flat)
        out, o = [], 0
        for t in tensors:
            out.append(flat[o:o + t.numel()].reshape(t.shape))
            o += t.numel()
        return out
