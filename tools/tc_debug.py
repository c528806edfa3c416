"""First-contact check of the tensor-core state kernel against the scalar one (prints, does not assert).
Usage: python tools/tc_debug.py D H Hp gamma N [T ncut]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import bsc_problem, rel_err  # noqa: E402
from oracle.common import DictAnneal  # noqa: E402
from prosper_b200 import _lib  # noqa: E402
from prosper_b200.em.camodels.bsc_et import BSC_ET  # noqa: E402

D, H, Hp, g, N = (int(v) for v in sys.argv[1:6])
T = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
ncut = float(sys.argv[7]) if len(sys.argv) > 7 else 0.0
y, params, _ = bsc_problem(D, H, N, 3)
an = DictAnneal(T=T, Ncut_factor=ncut, anneal_prior=False)
res = {}
for mode in (1, 2):
    m = BSC_ET(D, H, Hp, g)
    m.engine.set_state_kernel(mode)
    m._bind({'y': y})
    eng = m.engine
    p = m._pack_params(dict(params))
    a = eng.anneal(an)
    t0 = time.time()
    lse = eng.log_denominators(a, p, None, _lib.PASS_SELECT).cpu().numpy().copy()
    torch.cuda.synchronize()
    print("mode", mode, "lse done %.3fs" % (time.time() - t0), lse[:3], flush=True)
    t0 = time.time()
    st = eng.m_step_stats(a, p, None, _lib.PASS_SELECT).cpu().numpy().copy()
    torch.cuda.synchronize()
    print("mode", mode, "stats done %.3fs" % (time.time() - t0), flush=True)
    res[mode] = (lse, st, eng.layout)
l1, s1, lay = res[1]
l2, s2, _ = res[2]
print("lse   max abs diff", np.abs(l1 - l2).max())
for nm, off, cnt in (("Wp", lay.off_Wp, (D + 1) * lay.ld_Wp), ("Wq", lay.off_Wq, H * lay.ld_Wq), ("scalars", lay.off_scalars, lay.n_scalars)):
    a1, a2 = s1[off:off + cnt], s2[off:off + cnt]
    print("%-8s rel err %.3e   (max |ref| %.3e)" % (nm, rel_err(a2, a1), np.abs(a1).max()))
print("scalars", s1[lay.off_scalars:lay.off_scalars + 4], s2[lay.off_scalars:lay.off_scalars + 4])
