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


@pytest.mark.parametrize("n,m", [(1, 1), (10, 25), (64, 30), (100, 77), (129, 65), (1000, 676)])
def test_spd_solve_right(lib, dev, n, m):
    rng = np.random.RandomState(n)
    lda = (n + 7) // 8 * 8
    X = rng.standard_normal((n + 20, n))
    A = X.T @ X + 0.1 * np.eye(n)
    Bm = rng.standard_normal((m, n))
    Ad = torch.zeros(n, lda, dtype=torch.float64, device=dev); Ad[:, :n] = torch.as_tensor(A)
    Bd = torch.zeros(m, lda, dtype=torch.float64, device=dev); Bd[:, :n] = torch.as_tensor(Bm)
    work = torch.empty(lib.pet_spd_solve_work_doubles(n, lda), dtype=torch.float64, device=dev)
    info = C.c_int32(-1)
    assert lib.pet_spd_solve_right(n, m, P(Ad), lda, P(Bd), lda, P(work), C.byref(info), stream()) == 0
    assert info.value == 0
    ref = np.linalg.solve(A, Bm.T).T
    assert rel_err(Bd[:, :n].cpu().numpy(), ref) < 1e-9


def test_spd_solve_dead_unit_gives_lstsq_answer(lib, dev):
    """A unit that never fires leaves a zero row/column in Wq; np.linalg.lstsq returns the
    minimum-norm solution (zero column) and so must the device solve (bsc_et.py:380)."""
    rng = np.random.RandomState(0)
    n, m = 40, 12
    X = rng.standard_normal((60, n)); X[:, 7] = 0
    A = X.T @ X
    Bm = rng.standard_normal((m, n)); Bm[:, 7] = 0
    Ad = torch.as_tensor(A).to(dev).contiguous(); Bd = torch.as_tensor(Bm).to(dev).contiguous()
    work = torch.empty(lib.pet_spd_solve_work_doubles(n, n), dtype=torch.float64, device=dev)
    info = C.c_int32(-1)
    assert lib.pet_spd_solve_right(n, m, P(Ad), n, P(Bd), n, P(work), C.byref(info), stream()) == 0
    assert info.value == 1
    ref = np.linalg.lstsq(A, Bm.T, rcond=-1)[0].T
    assert rel_err(Bd.cpu().numpy(), ref) < 1e-10


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 1), (1000, 1000), (1000, 337), (1 << 20, 12345)])
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
