"""State-kernel time against the size of the state space (gamma = 2..5 at H' = 12): ms per 303 104 datapoints."""
import sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle.common import DictAnneal
from prosper_b200.em.camodels.bsc_et import BSC_ET
N = 303104
D, H, Hp = 676, 1000, 12
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(5)
rng = np.random.RandomState(5)
Wgt = rng.standard_normal((D, H)); Wgt *= 10 / np.linalg.norm(Wgt, axis=0, keepdims=True)
Wg = torch.as_tensor(Wgt).to(dev)
s = (torch.rand((N, H), device=dev, generator=gen) < 2.0 / H).to(torch.float64)
yt = s @ Wg.T + torch.randn((N, D), dtype=torch.float64, device=dev, generator=gen)
W0 = (yt.mean(0)[:, None] + 0.25 * torch.randn((D, H), dtype=torch.float64, device=dev, generator=gen)).cpu().numpy()
params = {'W': W0, 'pi': 1. / H, 'sigma': 1.2}
an = DictAnneal(T=1.0, Ncut_factor=0.0, anneal_prior=False)
for g in (2, 3, 4, 5):
    m = BSC_ET(D, H, Hp, g)
    m.engine.set_state_kernel(2)
    out = []
    for it in range(3):
        m.engine.enable_timing(True)
        m._fused_step(an, dict(params), {'y': yt})
        torch.cuda.synchronize()
        st = m.engine.stage_times()
        m.engine.enable_timing(False)
        out.append(round(st['state_kernel']['ms'], 3))
    S = sum(__import__('math').comb(Hp, k) for k in range(2, g + 1))
    print("gamma", g, "states", S, "chunks", -(-S // 64), "state_kernel ms", out, flush=True)
    del m
