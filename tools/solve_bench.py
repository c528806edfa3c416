"""Time the blocked Cholesky solve X.A = B (n = H = 1000, m = D = 676) and its factor-only part."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prosper_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda', 0)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
n, m = 1000, 676
STACKED = '--separate' not in sys.argv   # B under A in one buffer, as the engine holds them
g = torch.Generator(device=dev); g.manual_seed(0)
R = torch.randn(n, 2 * n, dtype=torch.float64, device=dev, generator=g)
A0 = R @ R.T / n + torch.eye(n, dtype=torch.float64, device=dev)
B0 = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g)
work = torch.empty(lib.pet_spd_solve_work_doubles(n, n), dtype=torch.float64, device=dev)
dropped = C.c_int32(0)
for mm, name in ((0, "factor only"), (m, "factor + solve")):
    ts = []
    for rep in range(12):
        AB = torch.cat([A0, B0]) if STACKED else None
        A, B = (AB[:n], AB[n:]) if STACKED else (A0.clone(), B0.clone())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        rc = lib.pet_spd_solve_right(n, mm, P(A), n, P(B), n, P(work), C.byref(dropped), st)
        e1.record(); torch.cuda.synchronize()
        assert rc == 0
        ts.append(e0.elapsed_time(e1))
    print("%s: median %.3f ms (min %.3f)" % (name, sorted(ts)[len(ts) // 2], min(ts)))
X = B
print("residual", float((X @ A0 - B0).abs().max() / B0.abs().max()), "dropped", dropped.value)
