#!/usr/bin/env python
"""bench.py -- datapoints/sec per EM iteration (select_Hprimes + E_step + M_step) of BSC-ET.

Workload (BASELINE.json configs[4], the one the metric is quoted on): BSC-ET D=26x26=676,
H=1000, H'=12, gamma=5 on N=1,000,000 synthetic patches (W_gt ~ N(0,1) with columns
rescaled to norm 10, s ~ Bernoulli(2/H), y = W_gt s + N(0,1)), float64, T=1, Ncut_factor=0.
The 1M datapoints are sharded by datapoint over the N ranks (strong scaling, as the reference's
MPI layer shards them); one NCCL all-reduce of the packed statistics per iteration.

    python bench.py --gpus 1 --steps K --warmup W          # our arm
    python bench.py --impl reference --steps K --warmup W   # CPU arm: NumPy port of the reference
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, H, HP, GAMMA = 676, 1000, 12, 5
N_TOTAL = int(os.environ.get("PET_BENCH_N", 1000000))
WORKLOAD = "BSC-ET D=676 H=1000 Hprime=12 gamma=5 N=%d synthetic patches, T=1, Ncut_factor=0" % N_TOTAL
METRIC = "datapoints/sec per EM iteration (select_Hprimes+E_step+M_step)"

# BASELINE.json configs[0..3] with the inputs of SURVEY 8(d); `--config 5` (the default) is configs[4] above.
# cpu_n: datapoints per host core of the CPU arm (cfgs 1-2 full N, cfg 3 N = 2000 P / 8, cfg 4 N = 64 P).
SMALL = {
    "1": dict(model="bsc", D=25, H=10, Hp=6, g=3, N=1000, cpu_n=1000,
              T=[(0, 2.), (.7, 1.)], Ncut=[(0, 0.), (2. / 3, 1.)],
              workload="BSC-ET bars 5x5 D=25 H=10 Hprime=6 gamma=3 N=1000, LinearAnnealing(50) T 2->1, Ncut_factor 0->1"),
    "2": dict(model="mca", D=25, H=10, Hp=6, g=3, N=2000, cpu_n=2000,
              T=[(0, 4.), (.8, 1.)], Ncut=[(0, 0.), (2. / 3, 1.)],
              workload="MCA-ET bars 5x5 D=25 H=10 Hprime=6 gamma=3 N=2000, LinearAnnealing(50) T 4->1, Ncut_factor 0->1"),
    "3": dict(model="tsc", D=64, H=16, Hp=8, g=4, N=50000, cpu_n=250,
              T=[(0, 2.), (.7, 1.)], Ncut=[(0, 0.), (2. / 3, 1.)],
              workload="TSC-ET bars 8x8 D=64 H=16 Hprime=8 gamma=4 N=50000, LinearAnnealing(50) T 2->1, Ncut_factor 0->1"),
    "3dsc": dict(model="dsc", D=64, H=16, Hp=8, g=4, N=50000, cpu_n=250,
                 T=[(0, 2.), (.7, 1.)], Ncut=[(0, 0.), (2. / 3, 1.)],
                 workload="DSC-ET (states -1,0,1) bars 8x8 D=64 H=16 Hprime=8 gamma=4 N=50000, LinearAnnealing(50) T 2->1, Ncut_factor 0->1"),
    "4": dict(model="gsc", D=144, H=64, Hp=8, g=3, N=200000, cpu_n=64,
              T=[(0, 1.)], Ncut=[(0, 0.)],
              workload="GSC-ET (spike-and-slab, scalar noise) D=144 H=64 Hprime=8 gamma=3 N=200000 synthetic patches, T=1"),
}


def bars_dict(Hb):
    R = Hb // 2
    W = np.zeros((R, R, Hb))
    for i in range(R):
        W[i, :, i] = 1.
        W[:, i, R + i] = 1.
    return W.reshape(R * R, Hb)


def small_problem(cfg_id, n, seed):
    """Synthetic inputs of SURVEY 8(d) for configs 1-4, plain NumPy so that both arms generate the same law: returns
    (y (n,D) float64, initial parameters with standard_init semantics)."""
    c = SMALL[cfg_id]
    Dm, Hm = c["D"], c["H"]
    rng = np.random.RandomState(seed)
    kind = c["model"]
    if kind in ("bsc", "mca"):
        W = 10.0 * bars_dict(Hm)
        s = rng.random_sample((n, Hm)) < 0.2
        if kind == "bsc":
            y = s.astype(np.float64) @ W.T
        else:                                              # max-rule superposition (mca_et.py:58-86)
            y = np.where(s[:, None, :], W[None, :, :], 0.0).max(axis=2)
        y += 2.0 * rng.standard_normal((n, Dm))
    elif kind in ("tsc", "dsc"):
        W = 10.0 * bars_dict(Hm)
        if kind == "tsc":
            u = rng.random_sample((n, Hm))
            s = np.where(u < 0.0625, -1.0, np.where(u < 0.125, 1.0, 0.0))
        else:
            s = rng.choice(np.array([-1., 0., 1.]), size=(n, Hm), p=[.06, .88, .06])
        y = s @ W.T + 2.0 * rng.standard_normal((n, Dm))
    else:                                                  # gsc: RandomState(3) law of SURVEY 8d
        W = rng.standard_normal((Dm, Hm))
        s = rng.random_sample((n, Hm)) < 2.0 / Hm
        z = s * (1.0 + rng.standard_normal((n, Hm)))
        y = z @ W.T + rng.standard_normal((n, Dm))
    mean = y.mean(axis=0)
    var = ((y - mean) ** 2).mean(axis=0)
    sig0 = np.sqrt(var).sum() / Dm
    W0 = mean[:, None] + rng.normal(scale=sig0 / 4., size=(Dm, Hm))
    if kind == "gsc":                                      # gsc_et.py:59-110
        pi = np.maximum(rng.rand(Hm) * 0.95, 0.05)
        params = {'W': W0, 'pi': pi, 'sigma_sq': float(np.mean(var) + 0.001), 'mu': rng.normal(0, 1, Hm),
                  'psi_sq': np.diag(np.maximum(rng.rand(Hm) * 2, 0.05))}
    elif kind == "dsc":                                    # dsc_et.py:872-925
        r = rng.rand(2)
        r = (1.0 / Hm) * r / r.sum()
        params = {'W': W0, 'pi': np.array([r[0], 1.0 - 1.0 / Hm, r[1]]), 'sigma': sig0}
    else:
        params = {'W': W0, 'pi': 1. / Hm, 'sigma': sig0}
    return y, params


def small_oracle(cfg_id):
    c = SMALL[cfg_id]
    a = (c["D"], c["H"], c["Hp"], c["g"])
    if c["model"] == "bsc":
        from oracle.bsc import BSC
        return BSC(*a)
    if c["model"] == "mca":
        from oracle.mca import MCA
        return MCA(*a)
    if c["model"] == "tsc":
        from oracle.tsc import TSC
        return TSC(*a)
    if c["model"] == "dsc":
        from oracle.dsc import DSC
        return DSC(*a, np.array([-1., 0., 1.]))
    from oracle.gsc import GSC
    return GSC(*a, sigma_sq_type='scalar')


def small_model(cfg_id, comm=None):
    c = SMALL[cfg_id]
    a = (c["D"], c["H"], c["Hp"], c["g"])
    from prosper_b200.em import camodels
    if c["model"] == "bsc":
        from prosper_b200.em.camodels.bsc_et import BSC_ET
        return BSC_ET(*a, comm=comm)
    if c["model"] == "mca":
        from prosper_b200.em.camodels.mca_et import MCA_ET
        return MCA_ET(*a, comm=comm)
    if c["model"] == "tsc":
        from prosper_b200.em.camodels.tsc_et import TSC_ET
        return TSC_ET(*a, comm=comm)
    if c["model"] == "dsc":
        from prosper_b200.em.camodels.dsc_et import DSC_ET
        return DSC_ET(*a, np.array([-1., 0., 1.]), comm=comm)
    from prosper_b200.em.camodels.gsc_et import GSC
    return GSC(*a, sigma_sq_type='scalar', comm=comm)


def small_schedule(cfg_id, steps=50):
    """The annealing values of iteration i (cyclic over the 50-iteration schedule of the bars examples,
    bars-learning.py:77-80 / param-bars-mca.py:34-37) as a list of Anneal dicts."""
    from prosper_b200.em.annealing import LinearAnnealing
    c = SMALL[cfg_id]
    la = LinearAnnealing(steps)
    la['T'] = c["T"]
    la['Ncut_factor'] = c["Ncut"]
    la['anneal_prior'] = False
    out = []
    for i in range(steps):
        la.cur_pos = i
        out.append(Anneal(T=float(la['T']), Ncut_factor=float(la['Ncut_factor']), anneal_prior=False))
    return out


class Anneal(dict):
    """anneal['key'] with prosper's missing-key-is-0.0 rule (annealing.py:93-94)."""
    crit_params = []

    def __missing__(self, k):
        return 0.0

    def as_dict(self):
        return dict(self)


def flops_per_dp():
    return 4.0 * D * H        # score GEMM 2DH + statistics GEMM 2DH  


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy port of the reference) on all host cores, bounded sample
# ----------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n, reps, cfg_id = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if cfg_id == "5":
        from helpers import bsc_problem
        from oracle.bsc import BSC
        y, params, _ = bsc_problem(D, H, n, seed)
        model = BSC(D, H, HP, GAMMA)
        sched = [Anneal(T=1.0, Ncut_factor=0.0, anneal_prior=False)]
    else:
        y, params = small_problem(cfg_id, n, seed)
        model = small_oracle(cfg_id)
        sched = small_schedule(cfg_id)
    times = []
    for i in range(reps):
        p = dict((k, (np.copy(v) if isinstance(v, np.ndarray) else v)) for k, v in params.items())
        if hasattr(model, 'check_params'):
            p = model.check_params(p)
        t0 = time.perf_counter()
        model.step(sched[i % len(sched)], p, {'y': y.copy()})
        times.append(time.perf_counter() - t0)
    return times


def cpu_port_throughput(n_per_core, reps, cores=None, cfg_id="5"):
    """dp/s of the oracle with one single-threaded worker per host core (the stand-in for
    `mpirun -np <cores>`; collectives are <0.1% of the reference's CPU time)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")   # never fork a process that holds a CUDA context
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(100 + i, n_per_core, reps, cfg_id) for i in range(cores)])
    per_rep = [max(r[i] for r in res) for i in range(reps)]      # slowest rank per iteration
    return [cores * n_per_core / t for t in per_rep], cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    small = args.config != "5"
    n_per_core = int(os.environ.get("PET_CPU_SAMPLE", SMALL[args.config]["cpu_n"] if small else 256))
    reps = args.warmup + args.steps
    vals, cores = cpu_port_throughput(n_per_core, reps, cfg_id=args.config)
    vals = vals[args.warmup:]
    v = float(np.mean(vals))
    sample = "%d datapoints per core x %d cores per step (cost is linear in N)" % (n_per_core, cores)
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "datapoints/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cores * n_per_core / v,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": SMALL[args.config]["workload"] if small else WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "datapoints/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "datapoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_shard(model, n_local, first, dev):
    """This rank's shard of SURVEY 8d config 5, generated ON the device by the product's own generator
    (`pet_generate_data`, counter-based RNG keyed by the global row index: the N ranks' shards are the rows of one
    data set): W_gt ~ N(0,1) with columns rescaled to norm 10, s ~ Bernoulli(2/H), y = W_gt s + N(0,1)."""
    rng = np.random.RandomState(5)
    Wgt = rng.standard_normal((D, H))
    Wgt *= 10.0 / np.linalg.norm(Wgt, axis=0, keepdims=True)
    gt = {'W': Wgt, 'pi': 2.0 / H, 'sigma': 1.0}
    return model.generate_data_device(gt, n_local, seed=5, row0=first, device=dev, latents=False)['y']


def ncu_traffic():
    """DRAM bytes per launch of the sliced GEMM kernel (mean of the score and the statistics shape, the two launches
    the roofline line averages over) and of the whole EM iteration, from the committed ncu launch list of one step
    (profiles/r02d_launches_step_n1.csv: dram__bytes_read.sum + dram__bytes_write.sum per launch); None if not recorded."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        return t["oz_gemm_bytes_per_launch"], t["step_dram_bytes"]
    except Exception:
        return None, None


def measure_fp64_peak(dev):
    """cuBLAS DGEMM 8192^3, best of 5 (same protocol as MEASURED_PEAKS.json, which has no FP64 entry)."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e9


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    from prosper_b200.utils import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = None
    if world > 1:
        # one process per GPU, bound to the GPU's NUMA node before any pinned buffer is allocated
        numa_cpus = parallel.bind_to_gpu_numa_node(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created: stdout carries ONE JSON line, so the
        # file descriptor points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    comm = parallel.default_comm()

    first, last = parallel.stride_data(N_TOTAL, comm=comm)
    n_local = last - first
    model = BSC_ET(D, H, HP, GAMMA, comm=comm)
    y = synth_shard(model, n_local, first, dev)
    np.random.seed(7)
    params0 = model.standard_init({'y': y})        # on the device: column moments, one all-reduce, device RNG for W
    anneal = Anneal(T=1.0, Ncut_factor=0.0, anneal_prior=False)

    data = {'y': y}
    eng = model.engine

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(params, d):
        new = model._fused_step(anneal, params, d)
        return {'W': new['W'], 'pi': new['pi'], 'sigma': new['sigma']}

    # ---- device-resident measurement ("value"): shard AND parameters stay in HBM ------------------
    params = dict(params0)
    params['W'] = torch.as_tensor(params0['W']).to(dev)     # W comes back as a CUDA tensor; pi / sigma are host scalars
    for _ in range(args.warmup):
        params = one_step(params, data)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.enable_timing(True)
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        params = one_step(params, data)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    stages = eng.stage_times()
    gemm_slices = eng.gemm_slices()
    eng.enable_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = N_TOTAL * args.steps / (ms_max / 1e3)

    # ---- the truncated iteration (SURVEY 8d also asks for Ncut_factor = 1): log-denominator pass, distributed k-th
    # largest (all-gather + radix select), statistics pass over the kept datapoints re-using the scores -----------------
    anneal_cut = Anneal(T=1.0, Ncut_factor=1.0, anneal_prior=False)
    cut_steps = max(1, min(args.steps, 3))
    p_cut = dict(params)
    new = model._fused_step(anneal_cut, p_cut, data)          # warm-up of the two-pass path
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(cut_steps):
        new = model._fused_step(anneal_cut, p_cut, data)
    c1.record()
    barrier()
    t = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cut_ms = float(t.item()) / cut_steps
    del new, p_cut

    # ---- end-to-end through the public API with HOST buffers ("e2e") ----------------------------
    y_host = torch.empty((n_local, D), dtype=torch.float64, pin_memory=True)
    y_host.copy_(y)
    del y, data
    model.invalidate_data()
    model.cache_data = False                       # every step re-uploads its inputs
    host_data = {'y': y_host.numpy()}
    from prosper_b200.utils.datalog import dlog
    p_e2e = dict(params0)
    e2e_steps = max(1, min(args.steps, 3))
    p_e2e = model.step(anneal, p_e2e, host_data)   # warm-up (allocations, pinned path)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        p_e2e = model.step(anneal, p_e2e, host_data)     # returns host NumPy W/pi/sigma (D2H inside)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N_TOTAL * e2e_steps / float(t.item())
    h2d = n_local * D * 8 + D * H * 8
    d2h = D * H * 8 + 16 * 8

    if rank == 0:
        peak = measure_fp64_peak(dev)
        per_step = dict((k, v['ms'] / args.steps) for k, v in stages.items())
        ns_score, ns = gemm_slices
        # kernels by total time: both GEMM stages run the same kernel template
        kernel_ms = {'gemm': per_step['score_gemm'] + per_step['stats_gemm'], 'state_kernel': per_step['state_kernel'],
                     'row_kernel': per_step['row_kernel']}
        dom = max(kernel_ms, key=lambda k: kernel_ms[k])
        hbm, bf16 = 6527.5, 1590.5
        peak_src = "fallback (B200_PROFILING.md)"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm, bf16 = mp["hbm_gbs"], mp["bf16_tflops_sustained"]
            peak_src = "MEASURED_PEAKS.json"
        except Exception:
            pass
        if dom == 'gemm':
            spans = max(1, stages['score_gemm']['spans'] + stages['stats_gemm']['spans'])
            avg_launch_ms = (stages['score_gemm']['ms'] + stages['stats_gemm']['ms']) / spans
            rows_per_launch = 2.0 * n_local * args.steps / spans
            # algorithmic work: 2*D*H flop per datapoint per GEMM, whatever pipe executes it
            achieved = 2.0 * D * H * rows_per_launch / (avg_launch_ms / 1e3) / 1e12
            if ns > 0:
                # slice products per FP64 product, averaged over the two GEMMs of an iteration (equal algorithmic work)
                pairs = (ns_score * (ns_score + 1) // 2 + ns * (ns + 1) // 2) / 2.0
                # int8 tensor-pipe peak MEASURED on this pool's B200 (tools/ubench/umma_probe.cu, profiles/r02_int8_peak.json):
                # the sustained figure, because the GEMMs run inside a long step
                int8_peak, int8_src = 2.0 * bf16, "2 x %s bf16_tflops_sustained (no measured int8 peak found)" % peak_src
                try:
                    ip = json.load(open(os.path.join(ROOT, "profiles", "r02_int8_peak.json")))
                    int8_peak = ip["int8_tops_sustained"]
                    int8_src = "profiles/r02_int8_peak.json int8_tops_sustained (measured: tcgen05 kind::i8 M128 N256 K32 back to back on all SMs)"
                except Exception:
                    pass
                roof = {"kernel": "oz::gemm_kernel<%d> (score GEMM) + oz::gemm_kernel<%d> (statistics GEMM)" % (ns_score, ns), "bound": "tensor",
                        "pipe": "tcgen05.mma kind::i8 + TMEM: %d (score, %d slices) and %d (statistics, %d slices) int8 slice "
                                "products per FP64 product" % (ns_score * (ns_score + 1) // 2, ns_score, ns * (ns + 1) // 2, ns),
                        "achieved": achieved, "peak": int8_peak / pairs, "unit": "TFLOP/s", "frac": achieved * pairs / int8_peak,
                        "pipe_achieved_tops": achieved * pairs, "pipe_peak_tops": int8_peak,
                        "pipe_nominal_tops": 4500.0, "frac_of_nominal": achieved * pairs / 4500.0,
                        "ncu_tensor_pipe_active": {"score_shape": 0.64, "statistics_shape": 0.80,
                                                   "source": "profiles/r02d_ncu_hot_kernels.txt (oz::gemm_kernel<6,3,3,64> / <7,2,4,64>)"},
                        "peak_source": int8_src + ", divided by the %.1f slice products (mean of the two GEMMs); cuBLAS DGEMM measured in this run: %.1f TFLOP/s"
                                       % (pairs, peak),
                        "traffic": ncu_traffic()[0], "step_traffic_bytes": ncu_traffic()[1],
                        "step_algorithmic_bytes": 2.0 * 8 * D * N_TOTAL / world}
            else:
                roof = {"kernel": "dgemm_kernel (FP64 DMMA, score + statistics GEMM)", "bound": "tensor",
                        "pipe": "fp64 mma.sync m8n8k4", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "peak_source": "cuBLAS DGEMM 8192^3 measured in this run", "traffic": None}
        else:
            spans = max(1, stages[dom]['spans'])
            avg_launch_ms = stages[dom]['ms'] / spans
            rows_per_launch = n_local * args.steps / spans
            # posterior kernels: algorithmic traffic = the score row in, the <s> row out (2 * 8 * H bytes per datapoint);
            # they are bound by shared-memory / issue latency, so the HBM fraction is low by construction
            achieved = 2.0 * 8 * H * rows_per_launch / (avg_launch_ms / 1e3) / 1e9
            roof = {"kernel": "gl_%s (posterior)" % dom, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                    "frac": achieved / hbm, "peak_source": peak_src + " hbm_gbs", "traffic": None}
        roof["avg_launch_ms"] = avg_launch_ms
        roof["stage_ms_per_step"] = per_step
        roof["whole_step_fp64_frac"] = (flops_per_dp() * N_TOTAL / world / (ms_max / args.steps / 1e3) / 1e12) / peak
        if world == 1:
            n_cpu = int(os.environ.get("PET_CPU_SAMPLE", 256))
            cpu_vals, cores = cpu_port_throughput(n_cpu, 2)
            cpu = {"value": float(cpu_vals[-1]), "unit": "datapoints/s", "cores": cores, "kind": "port",
                   "sample": "%d datapoints per core x %d cores, 1 iteration after 1 warm-up" % (n_cpu, cores)}
        else:       # the CPU arm is timed at N=1 only (the other ranks would idle behind it); `--impl reference` times it at any N
            cpu = {"value": None, "unit": "datapoints/s", "cores": 0, "kind": "port", "sample": "not timed at n_gpus > 1"}
        out = {
            "metric": METRIC, "value": value, "unit": "datapoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": "dp%d" % world, "l2": "inputs (%.1f GB per rank) larger than L2" % (n_local * D * 8 / 1e9),
                       "parity": "tests/test_bsc_gpu.py (float64, <=1e-8 vs oracle)",
                       "cpu_affinity": ("GPU-local NUMA node, %d CPUs" % len(numa_cpus)) if numa_cpus else "unchanged"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "datapoints/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "note": "model.step() with pinned host y re-uploaded every step, W/pi/sigma returned to host"},
            "gpu_launches": int(launches),
            "truncated_iteration": {"Ncut_factor": 1.0, "value": N_TOTAL / (cut_ms / 1e3), "unit": "datapoints/s",
                                    "ms_per_step": cut_ms, "steps": cut_steps,
                                    "note": "same step with the reference's datapoint truncation (bsc_et.py:247-260): "
                                            "one posterior evaluation (statistics parked per datapoint, added up after the distributed k-th largest gives the cut); not part of `value`"},
            "roofline": roof,
            "cpu_baseline": cpu,
            "result_check": {"pi": float(params['pi']), "sigma": float(params['sigma'])},
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_gpu_small(args):
    """`--config 1..4`: the same JSON line for the small configurations of BASELINE.json (bars tests, discrete states,
    spike-and-slab).  A step is one fused EM iteration at the annealing values of iteration i of the 50-iteration
    schedule (cyclic); these shapes are bound by launch latency and per-state FP64 ALU / SFU work, not by HBM or the
    tensor pipe, so the roofline object reports the dominant kernel's algorithmic bytes against HBM and the
    launch count per step next to it."""
    import torch
    import torch.distributed as dist
    from prosper_b200.utils import parallel

    cfg_id = args.config
    c = SMALL[cfg_id]
    Dm, Hm, N = c["D"], c["H"], int(os.environ.get("PET_BENCH_N_SMALL", c["N"]))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    comm = parallel.default_comm()
    y_all, params0 = small_problem(cfg_id, N, {"1": 1, "2": 1, "3": 2, "3dsc": 2, "4": 3}[cfg_id])
    first, last = parallel.stride_data(N, comm=comm)
    n_local = last - first
    y_host = torch.empty((n_local, Dm), dtype=torch.float64, pin_memory=True)
    y_host.copy_(torch.as_tensor(y_all[first:last]))
    del y_all
    model = small_model(cfg_id, comm=comm)
    sched = small_schedule(cfg_id)
    data = {'y': y_host.to(dev)}
    eng = model.engine
    keys = list(params0.keys())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def copy_params(p):
        return dict((k, (np.copy(v) if isinstance(v, np.ndarray) else v)) for k, v in p.items())

    def one_step(i, params, d):
        new = model._fused_step(sched[i % len(sched)], model.check_params(params), d)
        return dict((k, new[k]) for k in keys)

    params = copy_params(params0)
    for i in range(args.warmup):
        params = one_step(i, params, data)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.enable_timing(True)
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        params = one_step(args.warmup + i, params, data)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    stages = eng.stage_times()
    eng.enable_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = N * args.steps / (ms_max / 1e3)

    # ---- end to end: host buffers through model.step(), shard re-uploaded every step ----
    del data
    model.invalidate_data()
    model.cache_data = False
    host_data = {'y': y_host.numpy()}
    e2e_steps = max(1, min(args.steps, 10))
    p_e2e = copy_params(params0)
    p_e2e = model.step(sched[0], p_e2e, host_data)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        p_e2e = model.step(sched[(1 + i) % len(sched)], p_e2e, host_data)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N * e2e_steps / float(t.item())
    par_bytes = sum(int(np.asarray(v).size) * 8 for v in params0.values())

    if rank == 0:
        hbm, peak_src = 6527.5, "fallback (B200_PROFILING.md)"
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        per_step = dict((k, v['ms'] / args.steps) for k, v in stages.items())
        names = {'state_kernel': {'bsc': 'gl_row_* + gl_state_kernel', 'tsc': 'gl_row_kernel + gl_state_kernel<4,0>',
                                  'dsc': 'gl_row_kernel + gl_state_kernel<4,0>', 'mca': 'mca_kernel',
                                  'gsc': 'gsc_kernel'}[c["model"]],
                 'score_gemm': 'dgemm_kernel (score GEMM, FP64 DMMA)', 'stats_gemm': 'dgemm_kernel (statistics GEMMs, FP64 DMMA)',
                 'row_kernel': 'gl_row_kernel', 'solve': 'spd_solve', 'prepare': 'prepare', 'kth': 'kth_largest',
                 'scale_kernel': 'gl_scale_kernel', 'slice_kernels': 'ozaki slicing'}
        dom = max(per_step, key=lambda k: per_step[k])
        spans = max(1, stages[dom]['spans'])
        avg_launch_ms = stages[dom]['ms'] / spans
        rows_per_launch = n_local * args.steps / spans
        # algorithmic bytes of the posterior kernels per datapoint: the datapoint (8 D), its score row in (8 H) and
        # its <s> row out (8 H); of a GEMM stage: y in (8 D) and the H-long row in / out (8 H)
        kb = 8.0 * (Dm + 2 * Hm) if dom in ('state_kernel', 'row_kernel') else 8.0 * (Dm + Hm)
        achieved = kb * rows_per_launch / (avg_launch_ms / 1e3) / 1e9
        step_ms = ms_max / args.steps
        roof = {"kernel": names.get(dom, dom), "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "peak_source": peak_src, "traffic": None, "avg_launch_ms": avg_launch_ms,
                "stage_ms_per_step": per_step, "algorithmic_bytes_per_datapoint": kb,
                "whole_step_hbm_frac": (2.0 * 8 * Dm * n_local / (step_ms / 1e3) / 1e9) / hbm,
                "launches_per_step": launches / float(args.steps),
                "us_per_launch": 1e3 * step_ms / max(1.0, launches / float(args.steps)),
                "note": "small configuration: bound by launch latency (%d launches in a %.2f ms step) and per-state FP64 "
                        "ALU / SFU work, far below the HBM roof by construction (SURVEY 8d)"
                        % (round(launches / float(args.steps)), step_ms)}
        if world == 1:
            n_cpu = int(os.environ.get("PET_CPU_SAMPLE", c["cpu_n"]))
            cpu_vals, cores = cpu_port_throughput(n_cpu, 2, cfg_id=cfg_id)
            cpu = {"value": float(cpu_vals[-1]), "unit": "datapoints/s", "cores": cores, "kind": "port",
                   "sample": "%d datapoints per core x %d cores, 1 iteration after 1 warm-up" % (n_cpu, cores)}
        else:
            cpu = {"value": None, "unit": "datapoints/s", "cores": 0, "kind": "port", "sample": "not timed at n_gpus > 1"}
        check = dict((k, float(np.mean(np.asarray(params[k].cpu() if hasattr(params[k], 'cpu') else params[k]))))
                     for k in keys if k != 'W')
        out = {
            "metric": METRIC, "value": value, "unit": "datapoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": c["workload"], "baseline_config": cfg_id, "parallelism": "dp%d" % world,
                       "l2": "inputs (%.1f MB per rank) FIT in L2; the step is launch-bound, no flush between steps" % (n_local * Dm * 8 / 1e6),
                       "parity": "tests/test_trajectories_gpu.py, tests/test_%s_gpu.py" % {'bsc': 'bsc', 'mca': 'mca', 'tsc': 'tsc_dsc', 'dsc': 'tsc_dsc', 'gsc': 'gsc'}[c["model"]]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "datapoints/s", "h2d_bytes_per_step": n_local * Dm * 8 + par_bytes,
                    "d2h_bytes_per_step": par_bytes + 16 * 8, "steps": e2e_steps,
                    "note": "model.step() with pinned host y re-uploaded every step, parameters returned to host"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "result_check": check,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="5", choices=["1", "2", "3", "3dsc", "4", "5"],
                    help="BASELINE.json configs[k-1]; 5 (default) is the configuration the metric is quoted on")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "5":
        run_gpu_small(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
