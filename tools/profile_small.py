"""A few fused EM steps of one of the small BASELINE configurations (bench.py --config 1..4) for ncu captures.
Usage: python tools/profile_small.py <config> [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

cfg = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = bench.SMALL[cfg]
y, params = bench.small_problem(cfg, c["N"], 1)
m = bench.small_model(cfg)
sched = bench.small_schedule(cfg)
data = {'y': torch.as_tensor(y).cuda()}
keys = list(params.keys())
for i in range(steps):
    new = m._fused_step(sched[i], m.check_params(params), data)
    params = dict((k, new[k]) for k in keys)
torch.cuda.synchronize()
print("ok", cfg, steps)
