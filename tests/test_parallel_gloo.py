"""CPU tests of the N>1 host logic with a world_size-2 gloo process group: the mpi4py-style
communicator, stride_data sharding, allsort, and that a 2-rank run reproduces the 1-rank result
(oracle models driven through the same communicator interface the CUDA models use)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleComm(object):
    """Adapter: what oracle models call (allreduce, allsort) on top of prosper_b200's TorchComm."""

    def __init__(self, comm):
        from prosper_b200.utils import parallel
        self.comm, self.parallel = comm, parallel
        self.rank, self.size = comm.rank, comm.size

    def allreduce(self, x):
        return self.comm.allreduce(x)

    def allsort(self, a):
        return self.parallel.allsort(a, comm=self.comm)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prosper_b200.utils import parallel
    from helpers import bsc_problem
    from oracle.bsc import BSC
    from oracle.common import DictAnneal
    comm = parallel.default_comm()
    res = {}
    assert comm.rank == rank and comm.size == world
    # scalar / array reductions, broadcast
    res['sum_int'] = comm.allreduce(rank + 1)
    res['sum_float'] = comm.allreduce(0.5 * (rank + 1))
    res['sum_arr'] = comm.allreduce(np.arange(4, dtype=np.float64) * (rank + 1))
    res['max'] = comm.allreduce_max(10 + rank)
    res['bcast'] = comm.bcast({'a': rank}, root=0)
    t = torch.full((3,), float(rank + 1), dtype=torch.float64)
    res['tensor'] = comm.allreduce_tensor_(t).numpy().copy()
    # uneven shards (N=11 over 2 ranks -> 6 + 5, parallel.py:67-84)
    N = 11
    first, last = parallel.stride_data(N, comm=comm)
    res['range'] = (first, last)
    allv = np.random.RandomState(0).standard_normal(N)
    res['allsort'] = parallel.allsort(allv[first:last], comm=comm)
    res['allmean'] = parallel.allmean(allv[first:last, None] * np.ones((1, 3)), axis=0, comm=comm)
    # N-D arrays, equal shapes on all ranks (the reference's own tests/utils/test_parallel.py cases)
    a2 = np.random.RandomState(10 + rank).uniform(size=(10, 2))
    res['allsort_ax0'] = parallel.allsort(a2, axis=0, kind='quicksort', comm=comm)
    res['allsort_last'] = parallel.allsort(a2, comm=comm)
    res['allargsort_ax0'] = parallel.allargsort(a2, axis=0, kind='quicksort', comm=comm)
    res['allsum'] = parallel.allsum(np.ones(10), comm=comm)
    # a sharded EM step equals the single-rank step on the concatenated data
    y, params, _ = bsc_problem(25, 10, 301, 1, bars=True, pi=0.2, sigma=2.0)
    f, l = parallel.stride_data(y.shape[0], comm=comm)
    an = DictAnneal(T=1.5, Ncut_factor=0.6, anneal_prior=False)
    m = BSC(25, 10, 6, 3, comm=_OracleComm(comm))
    new = m.step(an, dict(params), {'y': y[f:l].copy()})
    res['W'], res['pi'], res['sigma'], res['L'], res['N_use'] = new['W'], new['pi'], new['sigma'], m.log['L'], m.log['N_use']
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([res], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import bsc_problem, rel_err
    from oracle.bsc import BSC
    from oracle.common import DictAnneal
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npy" % i), allow_pickle=True)[0] for i in range(2)]
    for i in range(2):
        assert r[i]['sum_int'] == 3 and r[i]['sum_float'] == 1.5 and r[i]['max'] == 11
        assert np.array_equal(r[i]['sum_arr'], np.arange(4) * 3.0)
        assert r[i]['bcast'] == {'a': 0}
        assert np.array_equal(r[i]['tensor'], np.full(3, 3.0))
    assert r[0]['range'] == (0, 6) and r[1]['range'] == (6, 11)
    allv = np.random.RandomState(0).standard_normal(11)
    for i in range(2):
        assert np.array_equal(r[i]['allsort'], np.sort(allv))
        assert np.allclose(r[i]['allmean'], allv.mean())
    # N-D: local sort, rank-ordered flat gather viewed with the axis lengthened (MPI Allgather), sort again
    loc = [np.random.RandomState(10 + i).uniform(size=(10, 2)) for i in range(2)]
    flat0 = np.concatenate([np.sort(x, 0).ravel() for x in loc])
    flat1 = np.concatenate([np.sort(x, -1).ravel() for x in loc])
    flata = np.concatenate([np.argsort(x, 0).ravel() for x in loc])
    for i in range(2):
        assert np.array_equal(r[i]['allsort_ax0'], np.sort(flat0.reshape(20, 2), 0))
        assert r[i]['allsort_last'].shape == (10, 4) and np.array_equal(r[i]['allsort_last'], np.sort(flat1.reshape(10, 4), -1))
        assert np.array_equal(r[i]['allargsort_ax0'], np.argsort(flata.reshape(20, 2), 0))
        assert r[i]['allsum'] == 20.0
    # 2-rank EM step == 1-rank EM step
    y, params, _ = bsc_problem(25, 10, 301, 1, bars=True, pi=0.2, sigma=2.0)
    an = DictAnneal(T=1.5, Ncut_factor=0.6, anneal_prior=False)
    m = BSC(25, 10, 6, 3)
    want = m.step(an, dict(params), {'y': y.copy()})
    for i in range(2):
        assert rel_err(r[i]['W'], want['W']) < 1e-10
        assert abs(r[i]['pi'] - want['pi']) < 1e-12 and abs(r[i]['sigma'] - want['sigma']) < 1e-12
        assert abs(r[i]['L'] - m.log['L']) < 1e-10 and r[i]['N_use'] == m.log['N_use']


def test_numa_binding_is_a_no_op_without_a_gpu():
    """bind_to_gpu_numa_node must never raise or change the affinity when NVML / the GPU is absent."""
    import os
    from prosper_b200.utils import parallel
    before = os.sched_getaffinity(0)
    assert parallel.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


@pytest.mark.skipif(not os.path.isdir("/root/reference/prosper/tests/utils"), reason="reference tree not present (GPU box)")
def test_the_references_own_unit_tests_pass_on_this_package(tmp_path):
    """prosper/tests/utils/test_{parallel,barstest,autotable,tracing}.py -- the only tests the reference has --
    run UNMODIFIED against `prosper_b200.install_as_prosper()`."""
    import subprocess
    code = (
        "import sys, unittest, importlib.util\n"
        "sys.path.insert(0, %r)\n"
        "import prosper_b200; prosper_b200.install_as_prosper()\n"
        "bad = 0; ran = 0\n"
        "for name in ('test_parallel', 'test_barstest', 'test_autotable', 'test_tracing'):\n"
        "    spec = importlib.util.spec_from_file_location('ref_' + name, '/root/reference/prosper/tests/utils/%%s.py' %% name)\n"
        "    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)\n"
        "    r = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(mod))\n"
        "    ran += r.testsRun; bad += len(r.failures) + len(r.errors)\n"
        "print('REFTESTS', ran, bad)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "REFTESTS 14 0" in out.stdout, out.stdout[-500:] + out.stderr[-3000:]
