"""Accuracy and throughput of the int8-sliced tcgen05 score GEMM against torch fp64 / the DMMA kernel."""
import ctypes as C, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prosper_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
shapes = [(128, 64, 64), (256, 128, 128), (300, 100, 70), (1000, 1000, 676), (16384, 1000, 676)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in sys.argv[1].split('x'))]
for (M, N, K) in shapes:
    ld = (K + 1) // 2 * 2
    A = torch.randn(M, ld, dtype=torch.float64, device=dev) * torch.exp(2 * torch.randn(M, 1, dtype=torch.float64, device=dev))
    B = torch.randn(N, ld, dtype=torch.float64, device=dev)
    ldc = (N + 1) // 2 * 2
    ref = A[:, :K] @ B[:, :K].T
    bound = (A[:, :K].abs().amax(1, keepdim=True) * B[:, :K].abs().amax(1)[None, :]) * K
    for ns in (6, 7):
        Cc = torch.full((M, ldc), float('nan'), dtype=torch.float64, device=dev)
        rc = lib.pet_ozaki_gemm_kk(M, N, K, P(A), ld, P(B), ld, P(Cc), ldc, ns, 1, st)
        if rc != 0:
            print("ERR", lib.pet_last_error()); sys.exit(1)
        err = (Cc[:, :N] - ref).abs()
        print("%dx%dx%d ns=%d max abs err %.2e  rel to max|C| %.2e  rel to K*amax*bmax %.2e nan %d" % (
            M, N, K, ns, float(err.max()), float(err.max() / ref.abs().max()), float((err / bound).max()),
            int(torch.isnan(Cc[:, :N]).sum())), flush=True)
    if M >= 16384:
        for ns in (6, 7):
            lib.pet_ozaki_gemm_kk(M, N, K, P(A), ld, P(B), ld, P(Cc), ldc, ns, 50, st)
            ms = lib.pet_ozaki_last_ms()
            print("  ns=%d gemm only %.3f ms -> %.1f effective FP64 TFLOP/s" % (ns, ms, 2.0 * M * N * K / ms / 1e9))

# ---- reduction over rows (statistics shape) --------------------------------------------------------------
shapes = [(5, 3, 7), (100, 70, 300), (677, 1000, 1000), (677, 1000, 16384), (677, 1000, 16001)]
for (M, N, K) in shapes:
    lda, ldb, ldc = (M + 1) // 2 * 2, (N + 1) // 2 * 2, (N + 1) // 2 * 2
    A = torch.randn(K, lda, dtype=torch.float64, device=dev) * torch.exp(2 * torch.randn(1, lda, dtype=torch.float64, device=dev))
    B = torch.rand(K, ldb, dtype=torch.float64, device=dev) * torch.exp(4 * torch.randn(1, ldb, dtype=torch.float64, device=dev))
    ref = A[:, :M].T @ B[:, :N]
    bound = (A[:, :M].abs().amax(0)[:, None] * B[:, :N].abs().amax(0)[None, :]) * K
    for ns in (6, 7):
        Cc = torch.full((M, ldc), float('nan'), dtype=torch.float64, device=dev)
        rc = lib.pet_ozaki_gemm_mn(M, N, K, P(A), lda, P(B), ldb, P(Cc), ldc, ns, 1, st)
        if rc != 0:
            print("ERR", lib.pet_last_error()); sys.exit(1)
        err = (Cc[:, :N] - ref).abs()
        print("MN %dx%dx%d ns=%d max abs err %.2e  rel to max|C| %.2e  rel to K*amax*bmax %.2e nan %d" % (
            M, N, K, ns, float(err.max()), float(err.max() / ref.abs().max()), float((err / bound).max()),
            int(torch.isnan(Cc[:, :N]).sum())), flush=True)
    if K >= 16384:
        for ns in (6, 7):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            lib.pet_ozaki_gemm_mn(M, N, K, P(A), lda, P(B), ldb, P(Cc), ldc, ns, 1, st)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            lib.pet_ozaki_gemm_mn(M, N, K, P(A), lda, P(B), ldb, P(Cc), ldc, ns, 50, st)
            ms = lib.pet_ozaki_last_ms()
            print("  ns=%d gemm+reduce %.3f ms -> %.1f effective FP64 TFLOP/s (slicing both operands once: %.3f ms)" % (
                ns, ms, 2.0 * M * N * K / ms / 1e9, (t1 - t0) * 1e3 - ms))
