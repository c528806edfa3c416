"""Small fused BSC step at the north-star shape, for ncu captures (not a benchmark)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.common import DictAnneal  # noqa: E402  (only for the anneal stand-in)
from prosper_b200.em.camodels.bsc_et import BSC_ET  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
D, H, Hp, g = 676, 1000, 12, int(os.environ.get('PET_PROF_GAMMA', 5))
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(5)
rng = np.random.RandomState(5)
Wgt = rng.standard_normal((D, H)); Wgt *= 10 / np.linalg.norm(Wgt, axis=0, keepdims=True)
Wg = torch.as_tensor(Wgt).to(dev)
s = (torch.rand((N, H), device=dev, generator=gen) < 2.0 / H).to(torch.float64)
yt = s @ Wg.T + torch.randn((N, D), dtype=torch.float64, device=dev, generator=gen)
W0 = (yt.mean(0)[:, None] + 0.25 * torch.randn((D, H), dtype=torch.float64, device=dev, generator=gen)).cpu().numpy()
params = {'W': W0, 'pi': 1. / H, 'sigma': 1.2}
m = BSC_ET(D, H, Hp, g)
if os.environ.get('PET_PROF_FORCE_TC'):
    m.engine.set_state_kernel(2)
an = DictAnneal(T=1.0, Ncut_factor=0.0, anneal_prior=False)
for _ in range(reps):
    new = m._fused_step(an, dict(params), {'y': yt})
    params = {'W': new['W'], 'pi': new['pi'], 'sigma': new['sigma']}
torch.cuda.synchronize()
print("ok", new['pi'], new['sigma'])
