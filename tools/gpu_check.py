"""First-contact diagnostics on the GPU box: prints parity numbers for every building block
(does not assert; pytest -m gpu does).  Usage: python tools/gpu_check.py [quick]"""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from prosper_b200 import _lib  # noqa: E402
from helpers import bsc_problem, rel_err, cand_mismatch_gap  # noqa: E402
from oracle.bsc import BSC  # noqa: E402
from oracle.common import DictAnneal  # noqa: E402

lib = _lib.load()
dev = torch.device('cuda', 0)


def P(t):
    return C.c_void_p(t.data_ptr())


def section(name):
    print("\n==== %s ====" % name, flush=True)


def timed(fn, reps=5):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def check_gemm():
    section("dgemm_kk / dgemm_mn vs torch fp64")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (M, N, K) in [(64, 64, 16), (100, 10, 26), (1000, 1000, 676), (4096, 1000, 676), (257, 130, 38), (33, 7, 5)]:
        lda = (K + 7) // 8 * 8
        A = torch.randn(M, lda, dtype=torch.float64, device=dev)
        B = torch.randn(N, lda, dtype=torch.float64, device=dev)
        ldc = (N + 7) // 8 * 8
        Cc = torch.full((M, ldc), 7.0, dtype=torch.float64, device=dev)
        rc = lib.pet_dgemm_kk(M, N, K, P(A), lda, P(B), lda, P(Cc), ldc, 1.0, 0.0, st)
        torch.cuda.synchronize()
        ref = A[:, :K] @ B[:, :K].T
        print("kk", (M, N, K), "rc", rc, "relerr", rel_err(Cc[:, :N].cpu().numpy(), ref.cpu().numpy()),
              "pad untouched", bool((Cc[:, N:] == 7.0).all().item()))
    for (M, N, K) in [(26, 10, 1000), (677, 1000, 16384), (65, 17, 300), (677, 1000, 5000)]:
        lda = (M + 7) // 8 * 8
        ldb = (N + 7) // 8 * 8
        A = torch.randn(K, lda, dtype=torch.float64, device=dev)
        B = torch.randn(K, ldb, dtype=torch.float64, device=dev)
        Cc = torch.zeros((M, ldb), dtype=torch.float64, device=dev)
        splits = lib.pet_dgemm_mn(M, N, K, None, lda, None, ldb, None, ldb, 0, None, 0, st)
        work = torch.empty(max(1, splits * M * ldb), dtype=torch.float64, device=dev)
        rc = lib.pet_dgemm_mn(M, N, K, P(A), lda, P(B), ldb, P(Cc), ldb, 0, P(work), work.numel(), st)
        rc2 = lib.pet_dgemm_mn(M, N, K, P(A), lda, P(B), ldb, P(Cc), ldb, 1, P(work), work.numel(), st)
        torch.cuda.synchronize()
        ref = 2 * (A[:, :M].T @ B[:, :N])
        print("mn", (M, N, K), "splits", splits, "rc", rc, rc2, "relerr", rel_err(Cc[:, :N].cpu().numpy(), ref.cpu().numpy()),
              "pad zero", bool((Cc[:, N:] == 0).all().item()))


def bench_gemm():
    section("FP64 throughput: cuBLAS (torch.matmul) vs DMMA tiles")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = timed(lambda: torch.matmul(a, b), reps=5)
    print("cuBLAS f64 8192^3: %.2f ms  %.2f TFLOP/s" % (ms, 2 * n ** 3 / ms / 1e9))
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    ms = timed(lambda: lib.pet_dgemm_kk(n, n, n, P(a), n, P(b), n, P(c), n, 1.0, 0.0, st), reps=5)
    print("pet_dgemm_kk 8192^3: %.2f ms  %.2f TFLOP/s" % (ms, 2 * n ** 3 / ms / 1e9))
    del a, b, c
    M, N, K = 65536, 1000, 676
    Y = torch.randn(M, 680, dtype=torch.float64, device=dev)
    W = torch.randn(N, 680, dtype=torch.float64, device=dev)
    Cc = torch.empty(M, 1000, dtype=torch.float64, device=dev)
    ms = timed(lambda: lib.pet_dgemm_kk(M, N, K, P(Y), 680, P(W), 680, P(Cc), 1000, 1.0, 0.0, st))
    print("score GEMM shape (65536x1000x676): %.3f ms  %.2f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
    ms = timed(lambda: torch.matmul(Y[:, :K], W[:, :K].T))
    print("  cuBLAS same shape: %.3f ms  %.2f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))
    S = torch.randn(M, 1000, dtype=torch.float64, device=dev)
    Wp = torch.zeros(677, 1000, dtype=torch.float64, device=dev)
    for rows in (16384, 65536):
        splits = lib.pet_dgemm_mn(677, 1000, rows, None, 680, None, 1000, None, 1000, 0, None, 0, st)
        work = torch.empty(splits * 677 * 1000, dtype=torch.float64, device=dev)
        ms = timed(lambda: lib.pet_dgemm_mn(677, 1000, rows, P(Y), 680, P(S), 1000, P(Wp), 1000, 1, P(work), work.numel(), st))
        print("stats GEMM shape (677x1000x%d, splits %d): %.3f ms  %.2f TFLOP/s" % (rows, splits, ms, 2.0 * 677 * 1000 * rows / ms / 1e9))
        ms = timed(lambda: torch.matmul(Y[:rows, :677].T, S[:rows]))
        print("  cuBLAS same shape: %.3f ms  %.2f TFLOP/s" % (ms, 2.0 * 677 * 1000 * rows / ms / 1e9))


def check_solve():
    section("spd_solve_right vs numpy")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.RandomState(0)
    for (n, m) in [(10, 25), (64, 30), (100, 77), (1000, 676)]:
        lda = (n + 7) // 8 * 8
        X = rng.standard_normal((n + 20, n))
        A = X.T @ X + 0.1 * np.eye(n)
        Bm = rng.standard_normal((m, n))
        Ad = torch.zeros(n, lda, dtype=torch.float64, device=dev); Ad[:, :n] = torch.as_tensor(A)
        Bd = torch.zeros(m, lda, dtype=torch.float64, device=dev); Bd[:, :n] = torch.as_tensor(Bm)
        work = torch.empty(lib.pet_spd_solve_work_doubles(n, lda), dtype=torch.float64, device=dev)
        info = C.c_int32(-1)
        t0 = time.time()
        rc = lib.pet_spd_solve_right(n, m, P(Ad), lda, P(Bd), lda, P(work), C.byref(info), st)
        dt = time.time() - t0
        ref = np.linalg.solve(A, Bm.T).T
        print("n=%d m=%d rc=%d info=%d relerr=%.3e wall=%.2f ms" % (n, m, rc, info.value, rel_err(Bd[:, :n].cpu().numpy(), ref), dt * 1e3))
    # dead unit: zero row/col -> lstsq minimum-norm answer has a zero column there
    n, m = 40, 12
    lda = 40
    X = rng.standard_normal((60, n)); X[:, 7] = 0
    A = X.T @ X
    Bm = rng.standard_normal((m, n)); Bm[:, 7] = 0
    Ad = torch.as_tensor(A).to(dev).contiguous(); Bd = torch.as_tensor(Bm).to(dev).contiguous()
    work = torch.empty(lib.pet_spd_solve_work_doubles(n, lda), dtype=torch.float64, device=dev)
    info = C.c_int32(-1)
    rc = lib.pet_spd_solve_right(n, m, P(Ad), lda, P(Bd), lda, P(work), C.byref(info), st)
    ref = np.linalg.lstsq(A, Bm.T, rcond=-1)[0].T
    print("dead unit: rc=%d info=%d relerr=%.3e" % (rc, info.value, rel_err(Bd.cpu().numpy(), ref)))


def check_bsc(D, H, Hp, g, N, seed, bars=False, T=1.0, ncut=0.0, ap=False, tag=""):
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    y, params, gt = bsc_problem(D, H, N, seed, bars=bars, pi=(0.2 if bars else None), sigma=(2.0 if bars else 1.0))
    an = DictAnneal(T=T, Ncut_factor=ncut, anneal_prior=ap)
    o = BSC(D, H, Hp, g)
    od = {'y': y.copy()}
    p0 = dict((k, (v.copy() if hasattr(v, 'copy') else v)) for k, v in params.items())
    t0 = time.time()
    od = o.select_hprimes(p0, od)
    oss = o.e_step(an, p0, od)
    onew = o.m_step(an, p0, oss, od)
    t_or = time.time() - t0
    m = BSC_ET(D, H, Hp, g)
    p1 = dict((k, (v.copy() if hasattr(v, 'copy') else v)) for k, v in params.items())
    d = {'y': y.copy()}
    d = m.select_Hprimes(p1, d)
    bad, gap = cand_mismatch_gap(od['_sim'], od['candidates'], d['candidates'])
    order_ok = float((od['candidates'] == d['candidates']).mean())
    # E-step on the ORACLE candidates so that columns line up even if a tie flipped
    d['candidates'] = od['candidates'].copy()
    ss = m.E_step(an, p1, d)
    e_lpj = np.abs(ss['logpj'] - oss['logpj']).max()
    from prosper_b200.utils.datalog import dlog, Keep
    keep = dlog.set_handler('*', Keep)
    new = m.M_step(an, p1, {'logpj': oss['logpj']}, d)
    L_compat = keep.last('L'); Nuse_compat = keep.last('N_use')
    p2 = dict((k, (v.copy() if hasattr(v, 'copy') else v)) for k, v in params.items())
    new_f = m._fused_step(an, p2, {'y': y.copy()})
    L_fused = keep.last('L'); Nuse_fused = keep.last('N_use')
    dlog.remove_handler(keep)
    print("BSC%s D=%d H=%d H'=%d g=%d N=%d T=%g ncut=%g ap=%d | cand rows differing %d (max gap %.2e) order-eq %.4f | "
          "logpj abs %.2e | compat: W %.2e pi %.2e sig %.2e L %.2e Nuse %d/%d | fused: W %.2e pi %.2e sig %.2e L %.2e Nuse %d | pivots %s | oracle %.1fs"
          % (tag, D, H, Hp, g, N, T, ncut, ap, bad, gap, order_ok, e_lpj,
             rel_err(new['W'], onew['W']), abs(new['pi'] - onew['pi']) / onew['pi'], abs(new['sigma'] - onew['sigma']) / onew['sigma'],
             abs(L_compat - o.log['L']) / abs(o.log['L']), Nuse_compat, o.log['N_use'],
             rel_err(new_f['W'], onew['W']), abs(new_f['pi'] - onew['pi']) / onew['pi'], abs(new_f['sigma'] - onew['sigma']) / onew['sigma'],
             abs(L_fused - o.log['L']) / abs(o.log['L']), Nuse_fused, getattr(m, 'last_dropped_pivots', None), t_or), flush=True)
    return m


def check_kth():
    section("kth_largest vs numpy")
    from prosper_b200.em.camodels import Engine
    eng = Engine(_lib.MODEL_BSC, 25, 10, 6, 3)
    rng = np.random.RandomState(3)
    for n, k in [(1, 1), (1000, 1), (1000, 1000), (1000, 337), (1 << 20, 12345)]:
        v = rng.standard_normal(n) * 100
        if n >= 1000:
            v[:10] = v[10]        # ties
        t = torch.as_tensor(v).to(dev)
        out = eng.kth_largest(t, k).cpu().numpy()[0]
        print("n=%d k=%d got %.17g want %.17g %s" % (n, k, out, np.sort(v)[-k], out == np.sort(v)[-k]))


def bench_step():
    section("fused step timing, BSC north-star shape on one GPU")
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    D, H, Hp, g = 676, 1000, 12, 5
    for N in (65536, 262144):
        rng = np.random.RandomState(5)
        Wgt = rng.standard_normal((D, H)); Wgt *= 10 / np.linalg.norm(Wgt, axis=0, keepdims=True)
        yt = torch.empty((N, D), dtype=torch.float64, device=dev)
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        Wg = torch.as_tensor(Wgt).to(dev)
        for a0 in range(0, N, 65536):
            b0 = min(N, a0 + 65536)
            s = (torch.rand((b0 - a0, H), device=dev, generator=gen) < 2.0 / H).to(torch.float64)
            yt[a0:b0] = s @ Wg.T + torch.randn((b0 - a0, D), dtype=torch.float64, device=dev, generator=gen)
        W0 = (yt.mean(0)[:, None] + 0.25 * torch.randn((D, H), dtype=torch.float64, device=dev, generator=gen)).cpu().numpy()
        params = {'W': W0, 'pi': 1. / H, 'sigma': 1.2}
        m = BSC_ET(D, H, Hp, g)
        data = {'y': yt}
        for ncut in (0.0, 1.0):
            an = DictAnneal(T=1.0, Ncut_factor=ncut, anneal_prior=False)
            m._fused_step(an, dict(params), data)
            torch.cuda.synchronize()
            m.engine.enable_timing(True)
            t0 = time.time()
            reps = 3
            for _ in range(reps):
                new = m._fused_step(an, dict(params), data)
            torch.cuda.synchronize()
            dt = (time.time() - t0) / reps
            st = m.engine.stage_times()
            m.engine.enable_timing(False)
            print("N=%d ncut=%g: %.2f ms/iter -> %.3e dp/s | stages(ms/iter): %s | pi %.5f sigma %.4f" % (
                N, ncut, dt * 1e3, N / dt, {k: round(v['ms'] / reps, 3) for k, v in st.items()}, new['pi'], new['sigma']), flush=True)
        del m, yt


if __name__ == '__main__':
    quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
    print(torch.cuda.get_device_name(0), torch.__version__)
    steps = [check_gemm, check_solve, check_kth,
             lambda: (section("BSC parity vs oracle"),
                      check_bsc(25, 10, 6, 3, 1000, 1, bars=True, tag=" cfg1"),
                      check_bsc(25, 10, 6, 3, 1000, 1, bars=True, T=2.0, ncut=0.5, tag=" cfg1"),
                      check_bsc(25, 10, 6, 3, 1000, 1, bars=True, T=1.5, ncut=0.3, ap=True, tag=" cfg1"),
                      check_bsc(100, 50, 8, 3, 2000, 2, tag=" mid"),
                      check_bsc(100, 50, 8, 4, 2000, 2, T=1.3, ncut=1.0, tag=" mid"),
                      check_bsc(676, 1000, 12, 5, 192, 5, tag=" cfg5-shape"),
                      check_bsc(676, 1000, 12, 5, 192, 5, T=1.2, ncut=1.0, tag=" cfg5-shape")),
             bench_gemm, bench_step]
    for s in steps:
        try:
            s()
        except Exception:
            traceback.print_exc()
            sys.stdout.flush()
