"""CAModel: base class of the ET sparse-coding models, bound to the CUDA engine.

Interface follows prosper/em/camodels/__init__.py:53-375 (constructor attributes, `step`
order, `standard_init`, `select_partial_data`, `compute_lpj`).  What differs is where the
O(N) work happens: `select_Hprimes`, `E_step`, `M_step` and the fused `step` call the C ABI
(include/prosper_b200.h) on device-resident data.
"""
import ctypes as C
import itertools

import numpy as np
import torch

from .. import Model
from ... import _lib
from ...utils import parallel, tracing
from ...utils.datalog import dlog


def generate_state_matrix(Hprime, gamma):
    """Binary H'-vectors with 2..gamma ones -> (state_list, no_states, state_matrix, state_abs).

    Host-side twin of camodels/__init__.py:21-47 (the engine enumerates the same order on its
    own; tests check both against each other)."""
    sl = [np.array(s, dtype=np.int8) for g in range(2, gamma + 1)
          for s in itertools.combinations(range(Hprime), g)]
    sm = np.zeros((len(sl), Hprime), dtype=np.uint8)
    for i, s in enumerate(sl):
        sm[i, s] = 1
    return sl, len(sl), sm, sm.sum(axis=1)


def _ptr(t):
    """Device/host address of a torch tensor or NumPy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


class Engine(object):
    """Thin RAII wrapper of a `pet_engine*`."""

    def __init__(self, model_kind, D, H, Hprime, gamma, states=None, device=None, chunk_rows=0):
        if not torch.cuda.is_available():
            raise RuntimeError("prosper_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._states = None if states is None else np.ascontiguousarray(states, dtype=np.float64)
        cfg = _lib.Config(model_kind, self.device, D, H, Hprime, gamma,
                          0 if states is None else len(self._states),
                          None if states is None else self._states.ctypes.data_as(_lib.c_double_p), chunk_rows)
        h = C.c_void_p()
        _lib.check(self.lib.pet_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.D, self.H, self.Hprime, self.gamma = D, H, Hprime, gamma
        self.S = int(self.lib.pet_num_states(h))
        self.Cols = int(self.lib.pet_num_columns(h))
        lay = _lib.StatsLayout()
        _lib.check(self.lib.pet_stats_layout_get(h, C.byref(lay)))
        self.layout = lay
        self.tdev = torch.device('cuda', self.device)
        self.stats = torch.zeros(lay.total, dtype=torch.float64, device=self.tdev)
        self.cut = torch.zeros(1, dtype=torch.float64, device=self.tdev)
        self.n = 0
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.pet_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def state_matrix(self):
        out = np.empty((self.S, self.Hprime), dtype=np.float64)
        _lib.check(self.lib.pet_state_matrix(self.h, out.ctypes.data_as(_lib.c_double_p)))
        return out

    # -- data ---------------------------------------------------------------------------
    def set_data(self, y, transient=False):
        """y: (n,D) float64 NumPy array (host; pinned if it came from torch) or CUDA tensor.
        transient: the shard will be re-uploaded for every pass (upload bound): short chunks, fine-grained overlap."""
        _lib.check(self.lib.pet_set_chunk_target(self.h, (128 << 20) if (transient and not isinstance(y, torch.Tensor)) else 0))
        if isinstance(y, torch.Tensor):
            if y.dtype != torch.float64 or y.dim() != 2 or y.stride(1) != 1:
                y = y.to(torch.float64).contiguous()
            n, ld = y.shape[0], y.stride(0)
        else:
            y = np.ascontiguousarray(y, dtype=np.float64)
            n, ld = y.shape[0], y.shape[1]
        assert y.shape[1] == self.D
        self._keep = [y]                      # the async copy reads it after we return
        _lib.check(self.lib.pet_set_data(self.h, _ptr(y), n, ld, self.stream()))
        self.n = n
        self.lse = torch.empty(max(n, 1), dtype=torch.float64, device=self.tdev)

    # -- parameter marshalling ----------------------------------------------------------
    def params(self, W, pi, sigma, mu=None):
        if isinstance(W, torch.Tensor):
            Wc = W if (W.dtype == torch.float64 and W.stride(-1) == 1) else W.to(torch.float64).contiguous()
            ldW = Wc.stride(0)
        else:
            Wc = np.ascontiguousarray(W, dtype=np.float64)
            ldW = Wc.shape[1]
        pi_arr = np.atleast_1d(np.asarray(pi, dtype=np.float64)).copy()
        mu_c = None
        if mu is not None and np.any(np.asarray(mu) != 0):
            mu_c = np.ascontiguousarray(mu, dtype=np.float64)
        p = _lib.Params(_ptr(Wc), ldW, pi_arr.ctypes.data_as(_lib.c_double_p), len(pi_arr), float(sigma), _ptr(mu_c))
        p._keep = (Wc, pi_arr, mu_c)
        return p

    @staticmethod
    def anneal(anneal):
        return _lib.Anneal(float(anneal['T']), float(anneal['Ncut_factor']), 1 if anneal['anneal_prior'] else 0)

    # -- operators ----------------------------------------------------------------------
    def select(self, p):
        cand = np.empty((self.n, self.Hprime), dtype=np.int64)
        _lib.check(self.lib.pet_select_hprimes(self.h, C.byref(p), _ptr(cand), self.stream()))
        return cand

    def set_candidates(self, cand):
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        assert cand.shape == (self.n, self.Hprime)
        _lib.check(self.lib.pet_set_candidates(self.h, _ptr(cand), self.stream()))
        torch.cuda.current_stream(self.tdev).synchronize()

    def e_step(self, a, p):
        logpj = np.empty((self.n, self.Cols), dtype=np.float64)
        _lib.check(self.lib.pet_e_step(self.h, C.byref(a), C.byref(p), _ptr(logpj), self.Cols, self.stream()))
        return logpj

    def e_step_device(self, a, p):
        """E_step whose logpj (n,C) stays in device memory (inference path)."""
        logpj = torch.empty((max(self.n, 1), self.Cols), dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.pet_e_step(self.h, C.byref(a), C.byref(p), _ptr(logpj), self.Cols, self.stream()))
        return logpj[:self.n]

    def posterior_topk(self, logpj, topK, logprob, marginals):
        """(idx (n,topK) int32, p (n,topK), m (n,H) or None) of pet_posterior_topk; tensors on the device."""
        n = self.n
        idx = torch.empty((max(n, 1), topK), dtype=torch.int32, device=self.tdev)
        pr = torch.empty((max(n, 1), topK), dtype=torch.float64, device=self.tdev)
        m = torch.empty((max(n, 1), self.H), dtype=torch.float64, device=self.tdev) if marginals else None
        _lib.check(self.lib.pet_posterior_topk(self.h, _ptr(logpj), logpj.stride(0), int(topK), 1 if logprob else 0,
                                               _ptr(idx), _ptr(pr), _ptr(m), self.stream()))
        return idx[:n], pr[:n], (m[:n] if marginals else None)

    def log_denominators(self, a, p, logpj=None, flags=0):
        _lib.check(self.lib.pet_log_denominators(self.h, C.byref(a), C.byref(p), _ptr(logpj),
                                                 0 if logpj is None else logpj.shape[1], flags,
                                                 _ptr(self.lse), self.stream()))
        return self.lse[:self.n]

    def kth_largest(self, vals, k):
        _lib.check(self.lib.pet_kth_largest(self.h, _ptr(vals), vals.numel(), int(k), _ptr(self.cut), self.stream()))
        return self.cut

    def m_step_stats(self, a, p, logpj=None, flags=0, use_cut=False):
        _lib.check(self.lib.pet_m_step_stats(self.h, C.byref(a), C.byref(p), _ptr(logpj),
                                             0 if logpj is None else logpj.shape[1], flags,
                                             1 if use_cut else 0, _ptr(self.cut) if use_cut else None,
                                             _ptr(self.stats), self.stream()))
        return self.stats

    def data_sum(self, use_cut):
        """sum over the kept datapoints of the engine's (mu-shifted) copy of y; device tensor (D,)."""
        out = torch.zeros(self.D, dtype=torch.float64, device=self.tdev)
        _lib.check(self.lib.pet_data_sum(self.h, 1 if use_cut else 0, _ptr(self.cut) if use_cut else None, _ptr(out),
                                         self.stream()))
        return out

    def solve(self, p, stats):
        W_new = torch.empty((self.D, self.H), dtype=torch.float64, device=self.tdev)
        info = C.c_int32(0)
        _lib.check(self.lib.pet_m_step_solve(self.h, C.byref(p), _ptr(stats), _ptr(W_new), C.byref(info), self.stream()))
        return W_new, int(info.value)

    def solve_rank_deficient(self, stats, rcond, add_colsum_diag):
        """Minimum-norm solution X = Wp^T . pinv(Wq) for a numerically singular Wq (dead units, N < H).

        np.linalg.lstsq / pinv (bsc_et.py:380, tsc_et.py:493) cut singular values below rcond * sigma_max;
        a Cholesky with dropped pivots only agrees with that when the null space is axis-aligned.  This rare,
        data-independent O(H^3) case goes through a symmetric eigendecomposition on the device."""
        lay, H, D = self.layout, self.H, self.D
        Wq = stats[lay.off_Wq:lay.off_Wq + H * lay.ld_Wq].reshape(H, lay.ld_Wq)[:, :H].clone()
        A = stats[lay.off_Wp:lay.off_Wp + (D + 1) * lay.ld_Wp].reshape(D + 1, lay.ld_Wp)[:, :H]
        if add_colsum_diag:
            Wq += torch.diag(A[D])
        lam, V = torch.linalg.eigh(0.5 * (Wq + Wq.T))
        # an exactly singular direction comes back from eigh as an eigenvalue of a few eps * lambda_max whose size depends
        # on the summation order of the statistics (atomics): cut at 32 eps so that the null space is recognised on every
        # run (LAPACK's gelsd sees an exact zero there; values between eps and 32 eps of lambda_max are rounding noise)
        keep = lam.abs() > max(rcond, 32 * 2.220446049250313e-16) * lam.abs().max()
        Vk = V[:, keep]
        return (A[:D] @ Vk) / lam[keep] @ Vk.T

    def scalars(self, stats):
        lay = self.layout
        return stats[lay.off_scalars:lay.off_scalars + lay.n_scalars].cpu().numpy()

    def enable_timing(self, on=True):
        _lib.check(self.lib.pet_enable_timing(self.h, 1 if on else 0))

    def stage_times(self):
        ns = _lib.N_STAGES
        out = (C.c_double * (2 * ns))()
        _lib.check(self.lib.pet_stage_times_ms(self.h, out))
        names = ['prepare', 'score_gemm', 'state_kernel', 'stats_gemm', 'solve', 'kth', 'row_kernel', 'scale_kernel',
                 'slice_kernels']
        return dict((nm, {'ms': out[i], 'spans': int(out[ns + i])}) for i, nm in enumerate(names))

    def gemm_path(self):
        """0 = FP64 DMMA kernels, n > 0 = int8 tcgen05 kernels with n slices per operand."""
        return int(self.lib.pet_gemm_path(self.h))

    def gemm_slices(self):
        """(score GEMM, statistics GEMM) int8 slices per operand; (0, 0) on the FP64 DMMA path."""
        import ctypes as C
        a, b = C.c_int32(0), C.c_int32(0)
        _lib.check(self.lib.pet_gemm_slices(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_state_kernel(self, mode):
        """0 = automatic, 1 = scalar FP64 state kernel, 2 = int8 tensor-core state kernel (fused BSC path)."""
        _lib.check(self.lib.pet_set_state_kernel(self.h, int(mode)))

    def state_kernel_path(self):
        return int(self.lib.pet_state_kernel_path(self.h))

    def launch_count(self):
        return int(self.lib.pet_launch_count(self.h))


class CAModel(Model):
    """Base of the ET models (camodels/__init__.py:53-193)."""

    model_kind = None

    def __init__(self, D, H, Hprime, gamma, to_learn=['W', 'pi', 'sigma'], comm=None):
        Model.__init__(self, comm)
        self.to_learn = to_learn
        self.D, self.H, self.Hprime, self.gamma = D, H, Hprime, gamma
        assert Hprime <= H
        assert gamma <= Hprime
        tol = 1e-5
        self.noise_policy = {
            'W': (-np.inf, +np.inf, False),
            'pi': (tol, 1. - tol, False),
            'sigma': (0., +np.inf, False),
        }
        self.state_list, self.no_states, self.state_matrix, self.state_abs = generate_state_matrix(Hprime, gamma)
        self.cache_data = True          # keep the device copy of my_data['y'] between calls
        self._engine = None
        self._bound = None

    # -- engine / data binding ------------------------------------------------------------
    def _make_engine(self):
        return Engine(self.model_kind, self.D, self.H, self.Hprime, self.gamma)

    @property
    def engine(self):
        if self._engine is None:
            self._engine = self._make_engine()
        return self._engine

    def invalidate_data(self):
        """Forget the device copy of the data (call after mutating my_data['y'] in place)."""
        self._bound = None
        self._bound_ref = None

    @staticmethod
    def _data_key(y):
        """Identity of a data array for the device-copy cache.  The bound array itself is kept alive next to the
        key (`_bound_ref`), so its address cannot be handed to another array while the cache entry lives; a
        NumPy buffer refilled in place is caught by a small content fingerprint (<= 4096 strided samples: the
        first and last rows and a diagonal walk), a torch tensor by its version counter."""
        if isinstance(y, torch.Tensor):
            return ('t', y.data_ptr(), tuple(y.shape), tuple(y.stride()), y.dtype, y._version)
        n = y.shape[0]
        if y.size:
            flat = y.reshape(-1) if y.flags.c_contiguous else None
            if flat is not None:
                step = max(1, flat.size // 4096)
                fp = (float(flat[::step].sum()), float(y[0].sum()), float(y[n - 1].sum()))
            else:
                fp = (float(y[::max(1, n // 64)].sum()), float(y[0].sum()), float(y[n - 1].sum()))
        else:
            fp = (0.0, 0.0, 0.0)
        return ('n', y.ctypes.data, y.shape, y.strides, y.dtype.str, fp)

    def _bind(self, my_data):
        y = my_data['y']
        key = self._data_key(y)
        if (self.cache_data and self._bound == key and getattr(self, '_bound_ref', None) is y
                and self.engine.n == y.shape[0]):
            return False
        self.engine.set_data(y, transient=not self.cache_data)
        self._bound = key
        self._bound_ref = y                # keeps the host array (and hence its address) alive while it is cached
        return True

    # -- reference interface --------------------------------------------------------------
    def generate_data(self, model_params, my_N):
        """camodels/__init__.py:104-122: Bernoulli(pi) latents, then generate_from_hidden."""
        s = np.random.random(size=(my_N, self.H)) < model_params['pi']
        return self.generate_from_hidden(model_params, {'s': s})

    def select_partial_data(self, anneal, my_data):
        """camodels/__init__.py:125-152: a random subset of ceil(partial * my_N) datapoints, in ascending order.
        A shard that lives on the device (CUDA tensors) is subsampled there: the permutation comes from the device RNG
        (seeded from np.random, so runs stay reproducible; SURVEY 8 f3: parity with np.random's stream is not required)
        and the rows are gathered by `pet_gather_rows` -- nothing returns to the host."""
        partial = anneal['partial']
        if partial == 0 or partial == 1:
            return my_data
        my_N = my_data['y'].shape[0]
        my_pN = int(np.ceil(my_N * partial))
        if my_N == my_pN:
            return my_data
        y = my_data['y']
        if isinstance(y, torch.Tensor) and y.is_cuda:
            gen = torch.Generator(device=y.device)
            gen.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))
            sel = torch.sort(torch.randperm(my_N, device=y.device, generator=gen)[:my_pN]).values
            lib = _lib.load()
            st = C.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)
            out = {}
            for k, v in my_data.items():
                if (isinstance(v, torch.Tensor) and v.is_cuda and v.dim() == 2 and v.dtype == torch.float64 and v.stride(1) == 1):
                    dst = torch.empty((my_pN, v.shape[1]), dtype=torch.float64, device=v.device)
                    _lib.check(lib.pet_gather_rows(my_pN, my_N, v.shape[1], _ptr(v), v.stride(0), _ptr(sel), _ptr(dst), dst.stride(0), st))
                    out[k] = dst
                elif isinstance(v, torch.Tensor):
                    out[k] = v[sel.to(v.device)]
                else:
                    out[k] = v[sel.cpu().numpy()]
            return out
        sel = np.random.permutation(my_N)[:my_pN]
        sel.sort()
        return dict((k, v[sel]) for k, v in my_data.items())

    def check_params(self, model_params):
        return model_params

    @tracing.traced
    def step(self, anneal, model_params, my_data):
        """One EM step in the order of camodels/__init__.py:163-193; the three starred calls
        of the reference are fused into `_fused_step` (logpj is never materialised)."""
        model_params = self.noisify_params(model_params, anneal)
        model_params = self.check_params(model_params)
        my_pdata = self.select_partial_data(anneal, my_data)
        new_model_params = self._fused_step(anneal, model_params, my_pdata)
        dlog.append_all(new_model_params)
        dlog.append_all(anneal.as_dict())
        return new_model_params

    def _fused_step(self, anneal, model_params, my_data):
        my_data = self.select_Hprimes(model_params, my_data)
        suff = self.E_step(anneal, model_params, my_data)
        return self.M_step(anneal, model_params, suff, my_data)

    def standard_init(self, data):
        """camodels/__init__.py:196-235; W noise is drawn on rank 0 and broadcast (the reference
        relies on identically seeded ranks)."""
        comm = self.comm
        my_y = data['y']
        if isinstance(my_y, torch.Tensor) and my_y.is_cuda:
            return self._standard_init_device(my_y)
        if isinstance(my_y, torch.Tensor):
            my_y = my_y.cpu().numpy()
        my_N, D = my_y.shape
        assert D == self.D
        W_mean = parallel.allmean(my_y, axis=0, comm=comm)
        sigma_sq = parallel.allmean((my_y - W_mean) ** 2, axis=0, comm=comm)
        sigma_init = np.sqrt(sigma_sq).sum() / D
        noise = np.random.normal(scale=sigma_init / 4., size=[D, self.H]) if comm.rank == 0 else None
        noise = comm.bcast(noise)
        return {'W': W_mean[:, None] + noise, 'pi': 1. / self.H, 'sigma': sigma_init}

    def _standard_init_device(self, y):
        """standard_init for a shard that lives on the device: column means and centred second
        moments with the engine's reduction kernels (two passes, as the reference's formula :217-220), one
        all-reduce each, W_init from the counter-based device RNG (same seed -> same W on every rank)."""
        import ctypes as C
        comm, D, H = self.comm, self.D, self.H
        lib = _lib.load()
        assert y.dim() == 2 and y.shape[1] == D and y.dtype == torch.float64 and y.stride(1) == 1
        st = C.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)
        n, ld = y.shape[0], y.stride(0)
        acc = torch.zeros(D + 1, dtype=torch.float64, device=y.device)
        _lib.check(lib.pet_colsum(n, D, _ptr(y), ld, _ptr(acc), st))
        acc[D] = float(n)
        comm.allreduce_tensor_(acc)
        N = float(acc[D].item())
        W_mean = (acc[:D] / N).contiguous()
        ssq = torch.zeros(D, dtype=torch.float64, device=y.device)
        _lib.check(lib.pet_col_centered_sumsq(n, D, _ptr(y), ld, _ptr(W_mean), _ptr(ssq), st))
        comm.allreduce_tensor_(ssq)
        sigma_init = float(torch.sqrt(ssq / N).sum().item()) / D
        seed = comm.bcast(int(np.random.randint(0, 2 ** 31 - 1)) if comm.rank == 0 else None)
        W = torch.empty((D, H), dtype=torch.float64, device=y.device)
        _lib.check(lib.pet_normal_fill(_ptr(W), H, D, H, _ptr(W_mean), sigma_init / 4., seed, st))
        return {'W': W.cpu().numpy(), 'pi': 1. / H, 'sigma': sigma_init}

    # -- data generation on the device -----------------------------------------
    def _latent_law(self, model_params):
        """(values, probabilities, combine) of one latent: Bernoulli(pi), linear superposition."""
        pi = float(model_params['pi'])
        return np.array([0., 1.]), np.array([1. - pi, pi]), 0

    def generate_data_device(self, model_params, my_N, seed=0, row0=0, device=None, latents=True):
        """`generate_data` (camodels/__init__.py:104-122) without the host: returns {'y': (my_N,D) float64 CUDA
        tensor, 's': (my_N,H) latent values (int8 if they are integers, else float64)}.  `row0` is the global
        index of the first datapoint: shards of one data set generated on different ranks (or in pieces) with
        the same seed equal the corresponding rows of one big call."""
        import ctypes as C
        lib = _lib.load()
        device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        values, probs, combine = self._latent_law(model_params)
        values = np.ascontiguousarray(values, dtype=np.float64)
        probs = np.ascontiguousarray(probs, dtype=np.float64)
        W = torch.as_tensor(np.ascontiguousarray(model_params['W'], dtype=np.float64)).to(device)
        assert W.shape == (self.D, self.H)
        y = torch.empty((my_N, self.D), dtype=torch.float64, device=device)
        sidx = torch.empty((my_N, self.H), dtype=torch.int8, device=device) if latents else None
        st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(lib.pet_generate_data(combine, my_N, row0, self.D, self.H, _ptr(W), self.H, len(values),
                                         values.ctypes.data_as(_lib.c_double_p), probs.ctypes.data_as(_lib.c_double_p),
                                         float(model_params['sigma']), int(seed), _ptr(y), self.D, _ptr(sidx), self.H, st))
        out = {'y': y}
        if latents:
            vals = torch.as_tensor(values).to(device)
            s = vals[sidx.long()]
            out['s'] = s.to(torch.int8) if np.all(values == np.round(values)) else s
        return out

    def compute_lpj(self, anneal, model_params, my_data):
        """camodels/__init__.py:238-253."""
        assert 'y' in my_data, "Key 'y' in my_data dict not defined."
        my_data = self.select_Hprimes(model_params, my_data)
        my_suff_stat = self.E_step(anneal, model_params, my_data)
        return my_suff_stat['logpj'], my_data['candidates']

    # -- inference (camodels/__init__.py:255-375) ---------------------------------------------
    _infer_block_rows = 32768        # datapoints per device block: logpj (rows x C) lives on the device only

    def _regenerate_states(self):
        self.state_list, self.no_states, self.state_matrix, self.state_abs = generate_state_matrix(self.Hprime, self.gamma)

    def _infer_res(self, my_N, topK):
        H = self.H
        return {'s': np.zeros((my_N, topK, H), dtype=np.int8), 'm': np.zeros((my_N, H)), 'p': np.zeros((my_N, topK)),
                'gamma': np.zeros((my_N,)), 'Hprime': np.zeros((my_N,))}

    def _infer_fill(self, res, rows, idx, p, m, cand, logpj, topK, logprob):
        """Write one block of results (rows = indices into res).  Base layout [null | h | states]:
        camodels/__init__.py:317-342.  Like the reference, entries of res['s'] left by an earlier
        adaptive round are overwritten, never cleared."""
        H = self.H
        res['p'][rows] = p.cpu().numpy()
        res['m'][rows] = m.cpu().numpy()                     # log domain until the end (:371)
        idx = idx.cpu().numpy().astype(np.int64)
        s = res['s']
        for k in range(topK):
            col = idx[:, k]
            single = (col >= 1) & (col < H + 1)
            s[rows[single], k, col[single] - 1] = 1
            multi = col >= H + 1
            if multi.any():
                sm = self.state_matrix[col[multi] - H - 1].astype(np.int8)          # (n_multi, Hprime)
                s[rows[multi][:, None], k, cand[multi]] = sm

    def _infer_logpj_device(self, a, model_params, y_block):
        """compute_lpj of one block with logpj left on the device -> (logpj (n,C) CUDA tensor, candidates (n,H') int64)."""
        block = self.select_Hprimes(model_params, {'y': y_block})
        cand = np.asarray(block['candidates']).astype(np.int64)
        return self.engine.e_step_device(a, self._pack_params(model_params)), cand

    def _infer_marginals(self):
        return True

    def _infer_kernel_logprob(self, logprob):
        return logprob

    def _infer_finish(self, res, logprob):
        if not logprob:
            res['m'] = np.exp(res['m'])                      # :371

    def _responsibilities(self, anneal, model_params, data):
        """mca_et.py:380-387 / dsc_et.py:776-784: row-normalised exp(logpj) of the compat E-step; the candidates are put
        back into ascending order first, as the reference does (in place)."""
        data['candidates'].sort(axis=1)
        F = torch.as_tensor(self.E_step(anneal, model_params, data)['logpj'])
        return torch.softmax(F, dim=1).cpu().numpy()

    def _infer_map_activity(self, res):
        return (res['s'][:, 0, :] != 0).sum(-1)              # :347

    def inference(self, anneal, model_params, test_data, topK=10, logprob=False, adaptive=True,
                  Hprime_max=None, gamma_max=None, **kwargs):
        """Top-K posterior states, their probabilities and the marginals of every cause
        (camodels/__init__.py:255-375; same arguments, same returned dict).  Candidate selection, the
        E-step, the row normalisation, the top-K search and the marginals run on the device
        (`pet_posterior_topk`); the adaptive H'/gamma growth loop stays on the host.  The engine's
        limits (H' <= 16, gamma <= 8) act as implicit Hprime_max / gamma_max."""
        assert 'y' in test_data, "Key 'y' in test_data dict not defined."
        model_params = self.check_params(model_params)
        comm = self.comm
        my_y = test_data['y']
        if isinstance(my_y, torch.Tensor):
            my_y = my_y.cpu().numpy()
        my_N, D = my_y.shape
        Hprime_start, gamma_start = self.Hprime, self.gamma
        hp_cap = min(self.H, _lib.MAX_HPRIME if Hprime_max is None else min(Hprime_max, _lib.MAX_HPRIME))
        g_cap = min(self.H, _lib.MAX_GAMMA if gamma_max is None else min(gamma_max, _lib.MAX_GAMMA))
        if topK == -1:
            topK = self.state_matrix.shape[0]
        res = self._infer_res(my_N, topK)
        self._infer_kwargs = kwargs
        which = np.ones(my_N, dtype=bool)
        saved_engine = self._engine
        try:
            while which.any():
                ind_n = np.where(which)[0]
                y_tmp = my_y[which]
                a = self.engine.anneal(anneal)
                for b0 in range(0, len(ind_n), self._infer_block_rows):
                    rows = ind_n[b0:b0 + self._infer_block_rows]
                    logpj, cand = self._infer_logpj_device(a, model_params, np.ascontiguousarray(y_tmp[b0:b0 + self._infer_block_rows]))
                    idx, pr, m = self.engine.posterior_topk(logpj, topK, self._infer_kernel_logprob(logprob),
                                                            self._infer_marginals())
                    res['Hprime'][rows] = self.Hprime
                    res['gamma'][rows] = self.gamma
                    self._infer_fill(res, rows, idx, pr, m, cand, logpj, topK, logprob)
                    del logpj
                if not adaptive:
                    break
                which = self._infer_map_activity(res) == self.gamma
                if not which.any():
                    break
                if self.Hprime >= hp_cap and self.gamma >= g_cap:
                    break
                print("Rank %i: For %i data points MAP state has activity equal to gamma." % (comm.rank, which.sum()))
                if self.Hprime < hp_cap:
                    self.Hprime += 1
                if self.gamma >= g_cap:
                    pass                                      # reference: `continue` without regenerating (:363-364)
                else:
                    self.gamma += 1
                print("Rank %i: Updating state matrix and running again." % comm.rank)
                self._regenerate_states()
                self._engine, self._bound = None, None        # engine for the grown (H', gamma)
        finally:
            self.Hprime, self.gamma = Hprime_start, gamma_start
            self._regenerate_states()
            # the training engine (if it was used at all) now holds the last inference block: forget the binding so
            # that the next step()/compute_lpj re-uploads its shard instead of silently running on the test data
            self._engine = saved_engine
            self.invalidate_data()
        self._infer_finish(res, logprob)
        comm.Barrier()
        return res

    # -- shared M-step plumbing -----------------------------------------------------------
    def _global_cut(self, lse, N_use_target):
        """k-th largest log-denominator over ALL ranks -> engine.cut (device scalar).

        Replaces parallel.allsort(all_denoms)[-N_use] (bsc_et.py:252): the log-denominators are
        all-gathered over NVLink (8 bytes per datapoint) and every rank runs the same radix
        select, exactly as every MPI rank sorts the gathered array in the reference."""
        comm, eng = self.comm, self.engine
        if comm.size == 1:
            return eng.kth_largest(lse, N_use_target)
        nmax = int(comm.allreduce_max(lse.numel()))
        pad = torch.full((nmax,), -float('inf'), dtype=torch.float64, device=lse.device)
        pad[:lse.numel()] = lse
        allv = torch.cat(comm.allgather_tensor(pad))
        return eng.kth_largest(allv, N_use_target)
