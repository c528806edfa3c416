"""Per-iteration differences of a 50-step trajectory (debug aid for tests/test_trajectories_gpu.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_trajectories_gpu as T  # noqa: E402
from helpers import rel_err  # noqa: E402
from oracle.common import DictAnneal  # noqa: E402

which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50


def run(m, o, anneal, y, params, keys, check=None):
    po, pm = T.cp(params), T.cp(params)
    it = 0
    while not anneal.finished and it < iters:
        an = DictAnneal(**anneal.as_dict())
        if check is not None:
            po = check(po)
        same = -1
        if which not in ('gsc', 'gscw'):
            od = o.select_hprimes(T.cp(po), {'y': y.copy()})
            md = m.select_Hprimes(m.check_params(T.cp(pm)), {'y': y.copy()})
            same = (np.sort(np.asarray(md['candidates']), 1) == np.sort(od['candidates'], 1)).all(1).mean()
        po = o.step(an, T.cp(po), {'y': y.copy()})
        pm = m.step(anneal, pm, {'y': y})
        anneal.next()
        errs = dict((k, rel_err(pm[k], po[k])) for k in keys)
        print(it, "T=%.3f ncut=%.2f" % (an['T'], an['Ncut_factor']), "same-cand-sets %.4f" % same,
              " ".join("%s %.2e" % kv for kv in errs.items()), getattr(m, '_inv_dropped', ''), flush=True)
        if max(errs.values()) > 1e-6:
            po = dict((k, (np.copy(np.asarray(pm[k])) if isinstance(pm[k], np.ndarray) else pm[k])) for k in po if k in pm)
        it += 1


import types
captured = {}


def fake_run(m, o, anneal, y, params, keys, tol=None, forced=False, inject_candidates=False):
    run(m, o, anneal, y, params, keys, o.check_params if inject_candidates else None)
    raise SystemExit(0)


T.run_trajectory = fake_run
if which == 'mca':
    T.test_mca_trajectory_cfg2()
elif which == 'gsc':
    T.test_gsc_trajectory_cfg4()
elif which == 'gscw':
    T.test_gsc_trajectory_well_conditioned()
else:
    T.test_tsc_dsc_trajectory_cfg3(which)
