"""Pinned host -> device bandwidth on this box (1-D and the pitched 2-D copy pet_set_data issues)."""
import torch, time, sys
n, D, ld = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 676, 680
h = torch.empty((n, D), dtype=torch.float64).pin_memory()
d = torch.empty((n, ld), dtype=torch.float64, device='cuda')
d1 = torch.empty((n, D), dtype=torch.float64, device='cuda')
side = torch.cuda.Stream()
def chunked(rows):
    for r0 in range(0, n, rows):
        d[r0:r0 + rows, :D].copy_(h[r0:r0 + rows], non_blocking=True)
def chunked1d(rows):
    for r0 in range(0, n, rows):
        d1[r0:r0 + rows].copy_(h[r0:r0 + rows], non_blocking=True)
cases = [("1-D whole", lambda: d1.copy_(h, non_blocking=True)), ("pitched 2-D whole", lambda: d[:, :D].copy_(h, non_blocking=True)),
         ("pitched 2-D chunks of 16384", lambda: chunked(16384)), ("pitched 2-D chunks of 131072", lambda: chunked(131072)),
         ("1-D chunks of 16384", lambda: chunked1d(16384))]
for name, fn in cases:
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s H2D: %.1f GB/s (%.1f ms)" % (name, h.numel() * 8 / dt / 1e9, dt * 1e3))
