"""N-GPU parity check, launched by tests/test_multi_gpu.py (or by hand) as
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py
Every rank takes its parallel.stride_data shard, runs fused EM steps with the NCCL all-reduce / all-gather,
and rank 0 compares with the single-rank oracle on the concatenated data."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import bsc_problem, rel_err  # noqa: E402
from oracle.bsc import BSC  # noqa: E402
from oracle.common import DictAnneal  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    from prosper_b200.em.camodels.mca_et import MCA_ET
    from prosper_b200.utils import parallel
    from prosper_b200.utils.datalog import dlog, Keep
    comm = parallel.default_comm()
    keep = dlog.set_handler('*', Keep)
    worst = 0.0
    for (D, H, Hp, g, N, T, ncut) in [(25, 10, 6, 3, 1001, 1.0, 0.0), (25, 10, 6, 3, 1001, 2.0, 0.7),
                                      (100, 50, 8, 4, 3000, 1.2, 1.0), (676, 1000, 12, 5, 333, 1.0, 1.0),
                                      (60, 40, 12, 5, 24000, 1.3, 1.0)]:      # tensor-core state kernel on every rank (forced below:
                                                                              # at 8 ranks 3 000 datapoints would pick the scalar one),
                                                                              # truncation with ONE posterior evaluation
        bars = D == 25
        y, params, _ = bsc_problem(D, H, N, 3, bars=bars, pi=(0.2 if bars else None), sigma=(2.0 if bars else 1.0))
        f, l = parallel.stride_data(N, comm=comm)
        an = DictAnneal(T=T, Ncut_factor=ncut, anneal_prior=False)
        m = BSC_ET(D, H, Hp, g, comm=comm)
        if N == 24000:
            m.engine.set_state_kernel(2)
        p = dict(params)
        po = dict(params)
        o = BSC(D, H, Hp, g)
        # (the N < H shape is run for one iteration only: its second-iteration Wq has ~270 singular values below
        #  machine precision, so W_new = pinv(Wq) Wp is ill-posed for ANY float64 implementation)
        for it in range(1 if N < H else 2):
            p = m._fused_step(an, dict(p), {'y': y[f:l].copy()})
            if comm.rank == 0:
                po = o.step(an, dict(po), {'y': y.copy()})
                errs = [rel_err(p['W'], po['W']), abs(p['pi'] - po['pi']) / po['pi'], abs(p['sigma'] - po['sigma']) / po['sigma'],
                        abs(keep.last('L') - o.log['L']) / abs(o.log['L'])]
                assert keep.last('N_use') == o.log['N_use'], (keep.last('N_use'), o.log['N_use'])
                # N < H: rank-deficient Wq, minimum-norm update (see tests/test_bsc_gpu.py): W to 1e-6, the rest to 1e-8
                if N < H:
                    assert errs[0] < 1e-6, errs
                    errs = errs[1:]
                worst = max(worst, max(errs))
                print("BSC D=%d H=%d N=%d ranks=%d it=%d: max rel err %.2e N_use %d" % (D, H, N, comm.size, it, max(errs), o.log['N_use']), flush=True)
    # every rank ends with identical parameters
    t = torch.as_tensor(p['W']).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref), "ranks diverged"
    if comm.rank == 0:
        assert worst < 1e-8, worst
        print("MULTI_GPU_OK worst=%.2e" % worst, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
