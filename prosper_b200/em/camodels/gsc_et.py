"""Gaussian Sparse Coding (spike-and-slab prior) with Expectation Truncation on the B200 engine.

Mirrors prosper/em/camodels/gsc_et.py (GSC): standard_init :59-110, check_params :161-191,
generate_data / generate_from_hidden :194-257, select_Hprimes :721-749 (+ component_scores :752-809),
E_step :401-580 (+ compute_posterior_hprime :260-398), M_step :584-718.

Device pipeline: Sigma^-1-weighted score GEMM and Gram matrix -> posterior kernel with k x k algebra per
state (`csrc/gsc_kernel.cu`) -> three statistics GEMMs + block scatter -> ONE all-reduce -> parameter
update (two H x H solves through the engine's Cholesky; the remaining O(H^2 D) arithmetic is
data-independent torch float64 on the device).  `sigma_sq_type='full'`: Sigma^-1 is formed on the host (D x D) and applied by
GEMMs (W^T Sigma^-1, y^T Sigma^-1 y); a symmetric Sigma is assumed.
"""
import ctypes as C

import numpy as np
import torch

from . import CAModel, Engine, _ptr
from ... import _lib
from ...utils import parallel
from ...utils import tracing

_SIGMA_TYPES = {'scalar': 0, 'diagonal': 1, 'full': 2}


class GSC(CAModel):
    model_kind = _lib.MODEL_GSC

    def __init__(self, D, H, Hprime=0, gamma=0, sigma_sq_type='scalar',
                 to_learn=['W', 'pi', 'mu', 'sigma_sq', 'psi_sq'], comm=None):
        CAModel.__init__(self, D, H, Hprime, gamma, to_learn, comm)
        tol = 1e-5
        self.noise_policy = {
            'W': (-np.inf, +np.inf, False), 'pi': (tol, 1. - tol, False), 'sigma_sq': (0., +np.inf, False),
            'mu': (-np.inf, +np.inf, False), 'psi_sq': (0., +np.inf, False),
        }
        if gamma <= 0 or gamma > H:              # gsc_et.py:45-50 (after the state matrix was built, as upstream)
            self.gamma = self.H
        if Hprime <= 0 or Hprime > H:
            self.Hprime = self.H
        elif Hprime < gamma:
            self.gamma = self.Hprime
        self._Hp_states, self._gamma_states = Hprime, gamma
        if Hprime <= 0 or gamma <= 0:
            # the reference's default Hprime=0 / gamma=0 means "no truncation": all 2^H states.  Only the truncated
            # posterior is built on the device (H' <= 16, gamma <= 8); say so here rather than on the first step.
            raise NotImplementedError("GSC(Hprime=%r, gamma=%r): the untruncated model (Hprime<=0 or gamma<=0, i.e. all "
                                      "2^H states) is not built on the device; pass 1 <= gamma <= Hprime <= %d"
                                      % (Hprime, gamma, _lib.MAX_HPRIME))
        self.sigma_sq_type = sigma_sq_type
        self.dtype_precision = np.float64

    def _make_engine(self):
        if self._Hp_states <= 0 or self._gamma_states <= 0:
            raise NotImplementedError("GSC with Hprime<=0 / gamma<=0 (no truncation) is not built on the device")
        return Engine(self.model_kind, self.D, self.H, self._Hp_states, self._gamma_states)

    # -- host-side pieces mirrored from the reference -------------------------------------------
    def standard_init(self, my_data):
        """gsc_et.py:59-110."""
        comm = self.comm
        temp = CAModel.standard_init(self, my_data)
        mp = {'W': temp['W'].copy()}
        pi = comm.bcast(np.random.rand(self.H)) * 0.95
        pi[pi < 0.05] = 0.05
        mp['pi'] = pi
        my_y = my_data['y']
        W_mean = parallel.allmean(my_y, axis=0, comm=comm)
        sigma_sq_sq = parallel.allmean((my_y - W_mean) ** 2, axis=0, comm=comm)
        if self.sigma_sq_type == 'full':
            mp['sigma_sq'] = np.diag(sigma_sq_sq) + (0.001 * np.eye(self.D))   # (upstream broadcasts a vector here)
        elif self.sigma_sq_type == 'diagonal':
            mp['sigma_sq'] = sigma_sq_sq + 0.001
        else:
            mp['sigma_sq'] = np.mean(sigma_sq_sq) + 0.001
        mp['mu'] = comm.bcast(np.random.normal(0, 1, [self.H])) if 'mu' in self.to_learn else np.zeros(self.H)
        if 'psi_sq' in self.to_learn:
            d = comm.bcast(np.random.rand(self.H)) * 2
            d[d < 0.05] = 0.05
            mp['psi_sq'] = np.diag(d)
        else:
            mp['psi_sq'] = np.eye(self.H)
        return comm.bcast(mp)

    def check_params(self, model_params):
        """gsc_et.py:161-191."""
        if self.comm.rank == 0:
            for k in ('W', 'mu', 'pi', 'psi_sq', 'sigma_sq'):
                assert np.isfinite(model_params[k]).all()
            if self.sigma_sq_type == 'full':
                assert np.sum(np.diag(model_params['sigma_sq']) <= 0) == 0
            else:
                assert np.sum(np.asarray(model_params['sigma_sq']) <= 0) == 0
        return model_params

    def generate_data_device(self, *args, **kwargs):
        raise NotImplementedError("GSC draws continuous latents (gsc_et.py:150-200): use generate_data on the host")

    def generate_data(self, model_params, my_N):
        s = np.zeros((my_N, self.H), dtype=bool)
        for n in range(my_N):
            s[n] = np.random.random(self.H) <= model_params['pi']
        return self.generate_from_hidden(model_params, {'s': s})

    def generate_from_hidden(self, model_params, my_hdata):
        """gsc_et.py:216-257 (including its `np.sum(indices) == 0` skip)."""
        D, H = self.D, self.H
        s = my_hdata['s']
        my_N = s.shape[0]
        y = np.zeros((my_N, D))
        z = np.zeros((my_N, H))
        if self.sigma_sq_type == 'full':
            sd = np.sqrt(model_params['sigma_sq'].diagonal())
        elif self.sigma_sq_type == 'diagonal':
            sd = np.sqrt(model_params['sigma_sq'])
        else:
            sd = np.sqrt(model_params['sigma_sq']) * np.ones(D)
        for n in range(my_N):
            act = np.nonzero(s[n])[0]
            if np.sum(act) == 0:
                continue
            z_n = np.random.multivariate_normal(model_params['mu'][act], model_params['psi_sq'][np.ix_(act, act)], 1).flatten()
            z[n, act] = z_n
            y[n] = model_params['W'][:, act] @ z_n + sd * np.random.randn(D)
        return {'y': y, 's': s, 'z': z}

    # -- marshalling -----------------------------------------------------------------------------
    def _pack(self, mp):
        W = np.ascontiguousarray(mp['W'], dtype=np.float64)
        pi = np.ascontiguousarray(mp['pi'], dtype=np.float64)
        mu = np.ascontiguousarray(mp['mu'], dtype=np.float64)
        psi = np.ascontiguousarray(mp['psi_sq'], dtype=np.float64)
        sig = np.ascontiguousarray(np.atleast_1d(mp['sigma_sq']), dtype=np.float64)
        assert W.shape == (self.D, self.H) and pi.shape == (self.H,) and mu.shape == (self.H,) and psi.shape == (self.H, self.H)
        p = _lib.GSCParams(_ptr(W), W.shape[1], pi.ctypes.data_as(_lib.c_double_p), mu.ctypes.data_as(_lib.c_double_p),
                           psi.ctypes.data_as(_lib.c_double_p), sig.ctypes.data_as(_lib.c_double_p),
                           _SIGMA_TYPES[self.sigma_sq_type])
        p._keep = (W, pi, mu, psi, sig)
        return p

    def _layout(self):
        if not hasattr(self, '_lay'):
            lay = _lib.GSCLayout()
            _lib.check(self.engine.lib.pet_gsc_layout_get(self.engine.h, C.byref(lay)))
            self._lay = lay
            self._stats = torch.zeros(lay.total, dtype=torch.float64, device=self.engine.tdev)
        return self._lay

    @staticmethod
    def cluster_order(cand):
        """Permutation grouping equal candidate sets, clusters in first-appearance order (dict order of
        gsc_et.py:733-745), original order inside a cluster."""
        _, first, inv = np.unique(cand, axis=0, return_index=True, return_inverse=True)
        rank_of_cluster = np.argsort(np.argsort(first))
        return np.argsort(rank_of_cluster[inv.ravel()], kind='stable')

    # -- compute_lpj / inference (gsc_et.py:811-944; inference itself is inherited, camodels/__init__.py:255-375) ----
    def _infer_logpj_device(self, a, model_params, y_block):
        eng = self.engine
        self._bind({'y': y_block})
        logpj = torch.empty((max(eng.n, 1), eng.Cols), dtype=torch.float64, device=eng.tdev)
        cand = np.empty((eng.n, self.Hprime), dtype=np.int64)
        _lib.check(eng.lib.pet_gsc_compute_lpj(eng.h, C.byref(self._pack(model_params)), _ptr(logpj), eng.Cols, _ptr(cand),
                                               eng.stream()))
        return logpj[:eng.n], cand

    def compute_lpj(self, anneal, model_params, my_data):
        """gsc_et.py:811-944 -> (logpj (n, 1+H+S), candidates (n,H')), rows in the order of my_data['y']."""
        assert 'y' in my_data, "Key 'y' in test_data dict not defined."
        y = my_data['y']
        logpj, cand = self._infer_logpj_device(None, model_params, y.cpu().numpy() if isinstance(y, torch.Tensor) else y)
        return logpj.cpu().numpy(), cand

    # -- the three operators ----------------------------------------------------------------------
    @tracing.traced
    def select_Hprimes(self, model_params, my_data):
        """gsc_et.py:721-749 -> my_data['data_clusters'] {key: {'hprimes','data','ind'}}."""
        eng = self.engine
        self._bind(my_data)
        cand = np.empty((eng.n, self.Hprime), dtype=np.int64)
        _lib.check(eng.lib.pet_gsc_select(eng.h, C.byref(self._pack(model_params)), _ptr(cand), eng.stream()))
        y = my_data['y']
        if isinstance(y, torch.Tensor):
            y = y.cpu().numpy()
        perm = self.cluster_order(cand)
        sorted_c = cand[perm]
        starts = np.flatnonzero(np.r_[True, np.any(sorted_c[1:] != sorted_c[:-1], axis=1)])
        ends = np.r_[starts[1:], len(perm)]
        clusters = {}
        for a, b in zip(starts, ends):
            idx = perm[a:b]
            clusters[str(sorted_c[a])] = {'hprimes': sorted_c[a].copy(), 'data': y[idx], 'ind': idx.tolist()}
        my_data['data_clusters'] = clusters
        self._cand = cand
        self._cand_key = self._bound
        return my_data

    @tracing.traced
    def E_step(self, anneal, model_params, my_data):
        """gsc_et.py:401-580 -> {'xpt_s','xpt_ss','xpt_sz','xpt_szsz'}; reorders my_data['y'] cluster-major and
        overwrites my_data['candidates'] (float), as the reference does (:572-573)."""
        eng = self.engine
        if getattr(self, '_cand', None) is None or self._cand_key != self._bound:
            raise RuntimeError("GSC.E_step needs select_Hprimes on the same data first")
        n, H = eng.n, self.H
        perm = self.cluster_order(self._cand)
        dst = np.empty(n, dtype=np.int64)
        dst[perm] = np.arange(n)
        dev = eng.tdev
        xs = torch.empty((n, H), dtype=torch.float64, device=dev)
        xsz = torch.empty((n, H), dtype=torch.float64, device=dev)
        xss = torch.empty((n, H, H), dtype=torch.float64, device=dev)
        xszsz = torch.empty((n, H, H), dtype=torch.float64, device=dev)
        _lib.check(eng.lib.pet_gsc_e_step(eng.h, C.byref(eng.anneal(anneal)), C.byref(self._pack(model_params)), _ptr(dst),
                                          _ptr(xs), _ptr(xss), _ptr(xsz), _ptr(xszsz), eng.stream()))
        y = my_data['y']
        y = y.cpu().numpy() if isinstance(y, torch.Tensor) else y
        my_data['y'] = y[perm]
        my_data['candidates'] = self._cand[perm].astype(np.float64)
        self.invalidate_data()
        self._cand = None
        return {'xpt_s': xs.cpu().numpy(), 'xpt_ss': xss.cpu().numpy(), 'xpt_sz': xsz.cpu().numpy(), 'xpt_szsz': xszsz.cpu().numpy()}

    @tracing.traced
    def M_step(self, anneal, model_params, suff_stats, my_data):
        """gsc_et.py:584-718 on caller-supplied moment tensors (compat path): the reductions over
        datapoints run through the engine's statistics GEMM / column-sum kernels."""
        eng = self.engine
        lay = self._layout()
        dev = eng.tdev
        H, D, ld = self.H, self.D, lay.ld
        y = my_data['y']
        yt = (y if isinstance(y, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(y, dtype=np.float64))).to(dev)
        n = yt.shape[0]
        st = self._stats
        st.zero_()
        pad = lambda t: torch.nn.functional.pad(t, (0, ld - H)).contiguous()
        xs = pad(torch.as_tensor(np.ascontiguousarray(suff_stats['xpt_s'], dtype=np.float64)).to(dev))
        xsz = pad(torch.as_tensor(np.ascontiguousarray(suff_stats['xpt_sz'], dtype=np.float64)).to(dev))
        xss = torch.as_tensor(np.ascontiguousarray(suff_stats['xpt_ss'], dtype=np.float64)).to(dev).reshape(n, H * H)
        xzz = torch.as_tensor(np.ascontiguousarray(suff_stats['xpt_szsz'], dtype=np.float64)).to(dev).reshape(n, H * H)
        ldy = (D + 1 + 7) // 8 * 8
        ya = torch.zeros((n, ldy), dtype=torch.float64, device=dev)
        ya[:, :D] = yt
        ya[:, D] = 1.0
        lib, s0 = eng.lib, eng.stream()

        def mn(M, N, A, lda, B, off):
            splits = lib.pet_dgemm_mn(M, N, n, None, lda, None, ld, None, ld, 0, None, 0, s0)
            work = torch.empty(max(1, splits * M * ld), dtype=torch.float64, device=dev)
            _lib.check(lib.pet_dgemm_mn(M, N, n, _ptr(A), lda, _ptr(B), ld, C.c_void_p(st.data_ptr() + 8 * off), ld, 0,
                                        _ptr(work), work.numel(), s0))
        mn(D + 1, H, ya, ldy, xsz, lay.off_A)
        mn(H, H, xs, ld, xsz, lay.off_Mssz)
        mn(H, H, xsz, ld, xsz, lay.off_Mout)
        tmp = torch.zeros(H * H, dtype=torch.float64, device=dev)
        for src, off, diag_off in ((xss, lay.off_ss, lay.off_sum_s), (xzz, lay.off_szsz, None)):
            tmp.zero_()
            _lib.check(lib.pet_colsum(n, H * H, _ptr(src), H * H, _ptr(tmp), s0))
            M = tmp.reshape(H, H).clone()
            if diag_off is not None:            # layout keeps diag(sum_ss) in sum_s
                st[diag_off:diag_off + H] = torch.diagonal(M)
                M.fill_diagonal_(0.0)
            st[off:off + H * ld].reshape(H, ld)[:, :H] = M
        _lib.check(lib.pet_colsum(n, D, _ptr(torch.square(ya[:, :D]).contiguous()), D, C.c_void_p(st.data_ptr() + 8 * lay.off_ysq), s0))
        if self.sigma_sq_type == 'full':        # sum_n y y^T (gsc_et.py:679-682)
            splits = lib.pet_dgemm_mn(D, D, n, None, ldy, None, ldy, None, lay.ld_yyT, 0, None, 0, s0)
            work = torch.empty(max(1, splits * D * lay.ld_yyT), dtype=torch.float64, device=dev)
            _lib.check(lib.pet_dgemm_mn(D, D, n, _ptr(ya), ldy, _ptr(ya), ldy, C.c_void_p(st.data_ptr() + 8 * lay.off_yyT), lay.ld_yyT, 0,
                                        _ptr(work), work.numel(), s0))
        st[lay.off_scalars] = float(n)
        return self._update(model_params, st)

    def _fused_step(self, anneal, model_params, my_data):
        eng = self.engine
        self._bind(my_data)
        lay = self._layout()
        _lib.check(eng.lib.pet_gsc_stats(eng.h, C.byref(eng.anneal(anneal)), C.byref(self._pack(model_params)),
                                         _lib.PASS_SELECT, _ptr(self._stats), eng.stream()))
        return self._update(model_params, self._stats)

    # -- parameter update from the packed statistics (gsc_et.py:622-716) ---------------------------
    def _inv(self, M):
        """inverse of a symmetric positive (semi-)definite H x H matrix via the engine's Cholesky solve"""
        eng = self.engine
        H = M.shape[0]
        ld = (H + 7) // 8 * 8
        A = torch.zeros((H, ld), dtype=torch.float64, device=M.device); A[:, :H] = M
        B = torch.zeros((H, ld), dtype=torch.float64, device=M.device); B[:, :H] = torch.eye(H, dtype=torch.float64, device=M.device)
        work = torch.empty(eng.lib.pet_spd_solve_work_doubles(H, ld), dtype=torch.float64, device=M.device)
        info = C.c_int32(0)
        _lib.check(eng.lib.pet_spd_solve_right(H, H, _ptr(A), ld, _ptr(B), ld, _ptr(work), C.byref(info), eng.stream()))
        self._inv_dropped = int(info.value)      # pivots the truncated Cholesky dropped (0: M was positive definite)
        return B[:, :H].clone()

    def _inv_szsz(self, sum_szsz, Wp, model_params, eps):
        """W_n = Wp . inv(sum <sz sz^T>) with the reference's fallbacks (gsc_et.py:623-637).  The reference calls
        np.linalg.inv (LU with partial pivoting), which only raises for an EXACTLY singular matrix and otherwise returns
        whatever the factorisation gives, however ill conditioned; the same factorisation is used here
        (torch.linalg.inv_ex on the device: a data-independent H x H operation) so that near-singular iterations follow
        the reference instead of a truncated Cholesky.  Singular -> pinv of the matrix plus an eps-sized rank-one
        perturbation -> if that fails too, the old W plus eps noise."""
        inv, info = torch.linalg.inv_ex(sum_szsz)
        self._inv_dropped = int(info.item())
        if self._inv_dropped == 0 and bool(torch.isfinite(inv).all()):
            return Wp @ inv
        dev, H = sum_szsz.device, self.H
        noise = self.comm.bcast(np.random.normal(0, eps, H) if self.comm.rank == 0 else None)
        noise = torch.as_tensor(np.outer(noise, noise), dtype=torch.float64, device=dev)
        try:
            pinv = torch.linalg.pinv(sum_szsz + noise, rtol=1e-15)
            if not bool(torch.isfinite(pinv).all()):
                raise RuntimeError("pinv of sum_szsz failed")
            return Wp @ pinv
        except RuntimeError:
            W_old = torch.as_tensor(np.asarray(model_params['W'], dtype=np.float64), device=dev)
            jitter = self.comm.bcast(np.random.normal(0, 1, [self.D, H]) if self.comm.rank == 0 else None)
            return W_old + eps * torch.as_tensor(jitter, dtype=torch.float64, device=dev)

    def _update(self, model_params, stats):
        comm = self.comm
        lay = self._layout()
        H, D, ld = self.H, self.D, lay.ld
        comm.allreduce_tensor_(stats)            # one collective: gsc_et.py:592,608-610,620,668,671,688/701/713
        N = int(round(float(stats[lay.off_scalars].item())))
        eps = 1e-5
        blk = lambda off, rows: stats[off:off + rows * ld].reshape(rows, ld)[:, :H]
        A = blk(lay.off_A, D + 1)
        Wp, sum_sz = A[:D], A[D]
        sum_s = stats[lay.off_sum_s:lay.off_sum_s + H]
        sum_ss = blk(lay.off_ss, H) + torch.diag(sum_s)
        sum_szsz = blk(lay.off_szsz, H) + torch.diag(stats[lay.off_sum_sz2:lay.off_sum_sz2 + H])
        M_ssz, M_out = blk(lay.off_Mssz, H), blk(lay.off_Mout, H)
        ysq = stats[lay.off_ysq:lay.off_ysq + D]
        W_n = self._inv_szsz(sum_szsz, Wp, model_params, eps)                 # :623-637
        if 'pi' in self.to_learn:                                             # :640-645
            model_params['pi'] = torch.clamp(sum_s / N, 5e-5, 1 - 5e-5).cpu().numpy()
        if 'W' in self.to_learn:
            model_params['W'] = W_n.cpu().numpy()
        if 'mu' in self.to_learn:                                             # :653-654
            model_params['mu'] = (sum_sz / (sum_s + np.finfo(np.float64).eps)).cpu().numpy()
        if 'psi_sq' in self.to_learn:                                         # :657-673
            mu = torch.as_tensor(model_params['mu'], dtype=torch.float64, device=stats.device)
            psi = torch.outer(mu, mu) * sum_ss + sum_szsz - 2 * (mu[:, None] * M_ssz)
            eye = torch.eye(H, dtype=torch.float64, device=stats.device)
            model_params['psi_sq'] = (psi * self._inv(sum_ss + eps * eye) + eps * eye).cpu().numpy()
        if 'sigma_sq' in self.to_learn:                                       # :675-715
            if self.sigma_sq_type == 'diagonal':
                model_params['sigma_sq'] = ((ysq - torch.einsum('dh,hk,dk->d', W_n, M_out, W_n)) / N + eps).cpu().numpy()
            elif self.sigma_sq_type == 'scalar':
                model_params['sigma_sq'] = float((ysq.sum() - torch.sum(M_out * (W_n.T @ W_n))) / N / D + eps)
            else:                                                               # :677-691
                yyT = stats[lay.off_yyT:lay.off_yyT + D * lay.ld_yyT].reshape(D, lay.ld_yyT)[:, :D]
                model_params['sigma_sq'] = ((yyT - W_n @ M_out @ W_n.T) / N
                                            + eps * torch.eye(D, dtype=torch.float64, device=stats.device)).cpu().numpy()
        return model_params
