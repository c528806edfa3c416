"""Where does the end-to-end step (host y re-uploaded every step) spend its time?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.common import DictAnneal
from prosper_b200.em.camodels.bsc_et import BSC_ET
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
D, H, Hp, g = 676, 1000, 12, 5
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(5)
y = torch.randn((N, D), dtype=torch.float64, device=dev, generator=gen)
yh = torch.empty((N, D), dtype=torch.float64, pin_memory=True); yh.copy_(y); del y
W0 = 0.25 * np.random.RandomState(1).standard_normal((D, H))
params = {'W': W0, 'pi': 1. / H, 'sigma': 1.2}
m = BSC_ET(D, H, Hp, g); m.cache_data = False
an = DictAnneal(T=1.0, Ncut_factor=0.0, anneal_prior=False)
data = {'y': yh.numpy()}
def sync(): torch.cuda.synchronize()
for rep in range(3):
    sync(); t0 = time.perf_counter()
    m.engine.set_data(data['y']); t1 = time.perf_counter()
    sync(); t2 = time.perf_counter()
    print("set_data call %.1f ms, upload done after %.1f ms (%.1f GB/s)" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, N * D * 8 / (t2 - t0) / 1e9))
for rep in range(3):
    sync(); t0 = time.perf_counter()
    new = m.step(an, dict(params), data); sync(); t1 = time.perf_counter()
    print("full step (host y) %.1f ms" % ((t1 - t0) * 1e3))
m.cache_data = True
m.step(an, dict(params), data)
for rep in range(2):
    sync(); t0 = time.perf_counter()
    new = m.step(an, dict(params), data); sync(); t1 = time.perf_counter()
    print("full step (cached shard) %.1f ms" % ((t1 - t0) * 1e3))
# upload-bound reference: select only (little compute) on freshly bound host data
m.cache_data = False
for rep in range(3):
    sync(); t0 = time.perf_counter()
    m.select_Hprimes(dict(params, mu=np.zeros(D)), {'y': data['y']}); sync(); t1 = time.perf_counter()
    print("select only (host y) %.1f ms" % ((t1 - t0) * 1e3))
