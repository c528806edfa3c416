"""Throughput of the two hot GEMM shapes for the tile variants (env PET_GEMM_KK / PET_GEMM_MN)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prosper_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


M, N, K = 16384, 1000, 676
Y = torch.randn(M, 680, dtype=torch.float64, device=dev); Y[:, 676] = 1; Y[:, 677:] = 0
W = torch.randn(N, 680, dtype=torch.float64, device=dev)
Cc = torch.empty(M, 1000, dtype=torch.float64, device=dev)
ms = timed(lambda: lib.pet_dgemm_kk(M, N, K, P(Y), 680, P(W), 680, P(Cc), 1000, 1.0, 0.0, st))
ref = Y[:, :K] @ W[:, :K].T
err = float((Cc - ref).abs().max() / ref.abs().max())
print("KK variant %s: score GEMM 16384x1000x676 %.3f ms %.2f TFLOP/s relerr %.1e" % (os.environ.get('PET_GEMM_KK', 'default'), ms, 2.0 * M * N * K / ms / 1e9, err))
S = torch.randn(M, 1000, dtype=torch.float64, device=dev)
Wp = torch.zeros(677, 1000, dtype=torch.float64, device=dev)
splits = lib.pet_dgemm_mn(677, 1000, M, None, 680, None, 1000, None, 1000, 0, None, 0, st)
work = torch.empty(splits * 677 * 1000, dtype=torch.float64, device=dev)
ms = timed(lambda: lib.pet_dgemm_mn(677, 1000, M, P(Y), 680, P(S), 1000, P(Wp), 1000, 0, P(work), work.numel(), st))
ref = Y[:, :677].T @ S
err = float((Wp - ref).abs().max() / ref.abs().max())
print("MN variant %s: stats GEMM 677x1000x16384 splits %d %.3f ms %.2f TFLOP/s relerr %.1e" % (os.environ.get('PET_GEMM_MN', 'default'), splits, ms, 2.0 * 677 * 1000 * M / ms / 1e9, err))
ms = timed(lambda: torch.matmul(Y[:, :K], W[:, :K].T)); print("cuBLAS score %.3f ms %.2f TF" % (ms, 2.0 * M * N * K / ms / 1e9))
ms = timed(lambda: torch.matmul(Y[:, :677].T, S)); print("cuBLAS stats %.3f ms %.2f TF" % (ms, 2.0 * 677 * 1000 * M / ms / 1e9))
