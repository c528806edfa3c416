"""GPU parity tests of the exported building blocks (through the C ABI)."""
import ctypes as C

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def P(t):
    return C.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch.device('cuda', 0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (64, 64, 16), (100, 10, 25), (33, 7, 5), (257, 130, 38),
                                   (1000, 1000, 676), (4099, 1000, 676)])
def test_dgemm_kk_matches_fp64_matmul(lib, dev, M, N, K):
    lda = (K + 7) // 8 * 8
    ldc = (N + 7) // 8 * 8
    g = torch.Generator(device=dev); g.manual_seed(M * 7 + N)
    A = torch.randn(M, lda, dtype=torch.float64, device=dev, generator=g)
    B = torch.randn(N, lda, dtype=torch.float64, device=dev, generator=g)
    Cm = torch.full((M, ldc), 3.0, dtype=torch.float64, device=dev)
    assert lib.pet_dgemm_kk(M, N, K, P(A), lda, P(B), lda, P(Cm), ldc, 1.0, 0.0, stream()) == 0
    ref = A[:, :K] @ B[:, :K].T
    assert rel_err(Cm[:, :N].cpu().numpy(), ref.cpu().numpy()) < 1e-13
    assert bool((Cm[:, N:] == 3.0).all())                       # padding untouched
    assert lib.pet_dgemm_kk(M, N, K, P(A), lda, P(B), lda, P(Cm), ldc, -0.5, 1.0, stream()) == 0
    assert rel_err(Cm[:, :N].cpu().numpy(), (0.5 * ref).cpu().numpy()) < 1e-13


@pytest.mark.parametrize("M,N,K", [(26, 10, 1000), (65, 17, 300), (677, 1000, 5000), (677, 1000, 16384), (9, 3, 1)])
def test_dgemm_mn_matches_fp64_matmul(lib, dev, M, N, K):
    lda = (M + 7) // 8 * 8
    ldb = (N + 7) // 8 * 8
    g = torch.Generator(device=dev); g.manual_seed(K)
    A = torch.randn(K, lda, dtype=torch.float64, device=dev, generator=g)
    B = torch.randn(K, ldb, dtype=torch.float64, device=dev, generator=g)
    Cm = torch.zeros((M, ldb), dtype=torch.float64, device=dev)
    splits = lib.pet_dgemm_mn(M, N, K, None, lda, None, ldb, None, ldb, 0, None, 0, stream())
    work = torch.empty(max(1, splits * M * ldb), dtype=torch.float64, device=dev)
    for acc in (0, 1):
        assert lib.pet_dgemm_mn(M, N, K, P(A), lda, P(B), ldb, P(Cm), ldb, acc, P(work), work.numel(), stream()) == 0
    ref = 2 * (A[:, :M].T @ B[:, :N])
    assert rel_err(Cm[:, :N].cpu().numpy(), ref.cpu().numpy()) < 1e-12
    assert bool((Cm[:, N:] == 0).all())


@pytest.mark.parametrize("stacked", [False, True])
@pytest.mark.parametrize("n,m", [(1, 1), (10, 25), (64, 30), (100, 77), (129, 65), (1000, 676)])
def test_spd_solve_right(lib, dev, n, m, stacked):
    """stacked: B directly under A in one buffer (what the engine does) -- the forward sweep is fused into the factorisation."""
    rng = np.random.RandomState(n)
    lda = (n + 7) // 8 * 8
    X = rng.standard_normal((n + 20, n))
    A = X.T @ X + 0.1 * np.eye(n)
    Bm = rng.standard_normal((m, n))
    if stacked:
        AB = torch.zeros(n + m, lda, dtype=torch.float64, device=dev)
        Ad, Bd = AB[:n], AB[n:]
    else:
        Ad = torch.zeros(n, lda, dtype=torch.float64, device=dev)
        Bd = torch.zeros(m, lda, dtype=torch.float64, device=dev)
    Ad[:, :n] = torch.as_tensor(A); Bd[:, :n] = torch.as_tensor(Bm)
    work = torch.empty(lib.pet_spd_solve_work_doubles(n, lda), dtype=torch.float64, device=dev)
    info = C.c_int32(-1)
    assert lib.pet_spd_solve_right(n, m, P(Ad), lda, P(Bd), lda, P(work), C.byref(info), stream()) == 0
    assert info.value == 0
    ref = np.linalg.solve(A, Bm.T).T
    assert rel_err(Bd[:, :n].cpu().numpy(), ref) < 1e-9


@pytest.mark.parametrize("stacked", [False, True])
def test_spd_solve_dead_unit_gives_lstsq_answer(lib, dev, stacked):
    """A unit that never fires leaves a zero row/column in Wq; np.linalg.lstsq returns the
    minimum-norm solution (zero column) and so must the device solve (bsc_et.py:380)."""
    rng = np.random.RandomState(0)
    n, m = 40, 12
    X = rng.standard_normal((60, n)); X[:, 7] = 0
    A = X.T @ X
    Bm = rng.standard_normal((m, n)); Bm[:, 7] = 0
    if stacked:
        AB = torch.as_tensor(np.concatenate([A, Bm])).to(dev).contiguous()
        Ad, Bd = AB[:n], AB[n:]
    else:
        Ad = torch.as_tensor(A).to(dev).contiguous(); Bd = torch.as_tensor(Bm).to(dev).contiguous()
    work = torch.empty(lib.pet_spd_solve_work_doubles(n, n), dtype=torch.float64, device=dev)
    info = C.c_int32(-1)
    assert lib.pet_spd_solve_right(n, m, P(Ad), n, P(Bd), n, P(work), C.byref(info), stream()) == 0
    assert info.value == 1
    ref = np.linalg.lstsq(A, Bm.T, rcond=-1)[0].T
    assert rel_err(Bd.cpu().numpy(), ref) < 1e-10


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 1), (1000, 1000), (1000, 337), (65536, 40000), (65537, 40000), (1 << 20, 12345)])
def test_kth_largest_bit_exact(dev, n, k):
    from prosper_b200 import _lib
    from prosper_b200.em.camodels import Engine
    eng = Engine(_lib.MODEL_BSC, 25, 10, 6, 3)
    rng = np.random.RandomState(3)
    v = rng.standard_normal(n) * 100
    if n >= 1000:
        v[:10] = v[10]            # ties
        v[20] = -0.0
        v[21] = 0.0
    got = eng.kth_largest(torch.as_tensor(v).to(dev), k).cpu().numpy()[0]
    assert got == np.sort(v)[-k]


# ---- FP64-accurate GEMMs on the int8 tcgen05 tensor cores (ozaki.cu) -----------------------------------
# Error model: every operand row is cut into ns signed 7-bit slices relative to its largest entry, slice
# products are exact, so |C - AB^T| <= ~K * 2^-(7 ns - 4) * rowmax(A) * rowmax(B); the tolerances below are
# that bound (worst case K = 1; for long dot products the error is ~100x smaller, 7 slices ~ FP64 rounding).
OZ_TOL = {6: 2e-12, 7: 2e-14}


@pytest.mark.parametrize("ns", [6, 7])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (128, 64, 64), (300, 100, 70), (33, 7, 5), (1000, 1000, 676), (4099, 1000, 676)])
def test_ozaki_gemm_kk_matches_fp64_matmul(lib, dev, M, N, K, ns):
    lda = (K + 1) // 2 * 2
    ldc = (N + 1) // 2 * 2
    g = torch.Generator(device=dev); g.manual_seed(M * 7 + N)
    A = torch.randn(M, lda, dtype=torch.float64, device=dev, generator=g)
    A *= torch.exp(3 * torch.randn(M, 1, dtype=torch.float64, device=dev, generator=g))       # row scales over ~5 decades
    B = torch.randn(N, lda, dtype=torch.float64, device=dev, generator=g)
    Cm = torch.full((M, ldc), 3.0, dtype=torch.float64, device=dev)
    assert lib.pet_ozaki_gemm_kk(M, N, K, P(A), lda, P(B), lda, P(Cm), ldc, ns, 1, stream()) == 0
    ref = A[:, :K] @ B[:, :K].T
    bound = K * A[:, :K].abs().amax(1, keepdim=True) * B[:, :K].abs().amax(1)[None, :]
    assert float(((Cm[:, :N] - ref).abs() / bound).max()) < OZ_TOL[ns]
    assert bool((Cm[:, N:] == 3.0).all())                       # padding untouched


@pytest.mark.parametrize("ns", [6, 7])
@pytest.mark.parametrize("M,N,K", [(5, 3, 7), (26, 10, 1000), (65, 17, 300), (677, 1000, 5000), (677, 1000, 16001)])
def test_ozaki_gemm_mn_matches_fp64_matmul(lib, dev, M, N, K, ns):
    lda, ldb = (M + 1) // 2 * 2, (N + 1) // 2 * 2
    g = torch.Generator(device=dev); g.manual_seed(K)
    A = torch.randn(K, lda, dtype=torch.float64, device=dev, generator=g)
    B = torch.rand(K, ldb, dtype=torch.float64, device=dev, generator=g)                      # posterior-like, in [0, 1)
    B *= torch.exp(4 * torch.randn(1, ldb, dtype=torch.float64, device=dev, generator=g))     # column scales
    Cm = torch.zeros((M, ldb), dtype=torch.float64, device=dev)
    assert lib.pet_ozaki_gemm_mn(M, N, K, P(A), lda, P(B), ldb, P(Cm), ldb, ns, 1, stream()) == 0
    ref = A[:, :M].T @ B[:, :N]
    bound = K * A[:, :M].abs().amax(0)[:, None] * B[:, :N].abs().amax(0)[None, :]
    assert float(((Cm[:, :N] - ref).abs() / bound).max()) < OZ_TOL[ns]


def test_ozaki_gemm_is_deterministic_and_exact_on_integers(lib, dev):
    """Small integers are represented exactly by the slices, so the product is exact; two runs agree bit for bit."""
    M, N, K = 257, 129, 200
    g = torch.Generator(device=dev); g.manual_seed(5)
    A = torch.randint(-60, 61, (M, K), device=dev, generator=g).to(torch.float64)
    B = torch.randint(-60, 61, (N, K), device=dev, generator=g).to(torch.float64)
    ldc = N + 1
    outs = []
    for _ in range(2):
        Cm = torch.zeros((M, ldc), dtype=torch.float64, device=dev)
        assert lib.pet_ozaki_gemm_kk(M, N, K, P(A), K, P(B), K, P(Cm), ldc, 7, 1, stream()) == 0
        outs.append(Cm[:, :N].clone())
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0], A @ B.T)


def test_ozaki_gemm_heavy_tailed_rows(lib, dev):
    """Rows whose entries span six decades (one dominant entry per row): the slices are cut relative to the LARGEST
    entry of a row, so the error is bounded by K 2^-45 rowmax(A) rowmax(B) -- relative to |C| that is only small when
    the dominant entries contribute to C, which is what this test pins down from both sides: the absolute bound always
    holds, and the error relative to |C| degrades by exactly the dynamic range when the dominant entry is multiplied by
    zero.  (The EM path slices datapoints and dictionary columns, whose entries share one scale)"""
    M, N, K = 512, 256, 676
    g = torch.Generator(device=dev); g.manual_seed(9)
    A = torch.randn(M, K, dtype=torch.float64, device=dev, generator=g)
    B = torch.randn(N, K, dtype=torch.float64, device=dev, generator=g)
    A[:, 0] *= 1e6                                              # the dominant column
    Cm = torch.zeros((M, N), dtype=torch.float64, device=dev)
    assert lib.pet_ozaki_gemm_kk(M, N, K, P(A), K, P(B), K, P(Cm), N, 7, 1, stream()) == 0
    ref = A @ B.T
    bound = K * A.abs().amax(1, keepdim=True) * B.abs().amax(1)[None, :]
    assert float(((Cm - ref).abs() / bound).max()) < OZ_TOL[7]                     # the model's bound
    assert float(((Cm - ref).abs() / ref.abs().clamp_min(1e-300)).median()) < 1e-12   # |C| is dominated by the big entries
    # the same rows against a B whose dominant-direction entries are zero: |C| ~ 1e-6 rowmax, the ABSOLUTE error stays
    B0 = B.clone(); B0[:, 0] = 0.0
    assert lib.pet_ozaki_gemm_kk(M, N, K, P(A), K, P(B0), K, P(Cm), N, 7, 1, stream()) == 0
    ref0 = A @ B0.T
    err0 = (Cm - ref0).abs()
    assert float((err0 / bound).max()) < OZ_TOL[7]
    rel0 = float((err0 / ref0.abs().clamp_min(1e-300)).median())
    assert 1e-12 < rel0 < 1e-6, rel0                            # ~2^-45 x 1e6: the documented loss, not more
