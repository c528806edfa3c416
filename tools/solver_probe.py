import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import bsc_problem, rel_err
from oracle.bsc import BSC
from oracle.common import DictAnneal
from prosper_b200.em.camodels.bsc_et import BSC_ET
D, H, Hp, g, N = 676, 1000, 12, 5, 333
y, params, _ = bsc_problem(D, H, N, 3)
an = DictAnneal(T=1.0, Ncut_factor=1.0, anneal_prior=False)
m = BSC_ET(D, H, Hp, g); o = BSC(D, H, Hp, g)
p = dict(params); po = dict(params)
for it in range(3):
    # same input params for both: isolates the per-step difference
    pin = dict(po)
    p = m._fused_step(an, dict(pin), {'y': y.copy()})
    # what does Wq look like?
    eng = m.engine; lay = eng.layout
    Wq = eng.stats[lay.off_Wq:lay.off_Wq + H * lay.ld_Wq].reshape(H, lay.ld_Wq)[:, :H].cpu().numpy().copy()
    Wq[np.arange(H), np.arange(H)] += eng.stats[lay.off_Wp + D * lay.ld_Wp: lay.off_Wp + D * lay.ld_Wp + H].cpu().numpy()
    sv = np.linalg.svd(Wq, compute_uv=False)
    po = o.step(an, dict(pin), {'y': y.copy()})
    rows = np.abs(p['W'] - po['W']).max(axis=0) / np.abs(po['W']).max()
    print("it", it, "relerr W", rel_err(p['W'], po['W']), "pivots dropped", m.last_dropped_pivots, "diag min/max", np.diag(Wq).min(), np.diag(Wq).max(),
          "sv min/max", sv.min(), sv.max(), "n sv < eps*max", (sv < 1.1e-16 * sv.max()).sum(), "bad cols", (rows > 1e-6).sum())
