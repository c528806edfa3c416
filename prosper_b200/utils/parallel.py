"""Data-parallel plumbing: the reference's mpi4py communicator, on torch.distributed.

Mirrors prosper/utils/parallel.py (stride_data :44-84, allsort :87-110, allmean :138-156,
allsum :159-170, pprint :27-41) and the subset of the mpi4py communicator API the models
call (`comm.rank`, `comm.size`, `comm.allreduce`, `comm.bcast`, `comm.Barrier`).  One process
per GPU; NCCL for device tensors, gloo for host tensors (CPU tests run with gloo only).
"""
import sys

import numpy as np
import torch
import torch.distributed as dist


class SerialComm(object):
    """Single-process communicator (what the reference sees without mpirun)."""
    rank = 0
    size = 1

    def allreduce(self, x):
        return x

    def allreduce_max(self, x):
        return x

    def bcast(self, x, root=0):
        return x

    def allreduce_tensor_(self, t):
        return t

    def allgather_tensor(self, t):
        return [t]

    def Barrier(self):
        pass


class TorchComm(object):
    """torch.distributed process group with the mpi4py-style calls the models use."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    def _host_device(self):
        # NCCL cannot reduce host tensors: stage small host values through the current GPU
        return torch.device('cuda', torch.cuda.current_device()) if self.backend == 'nccl' else torch.device('cpu')

    def allreduce(self, x):
        """Sum of a Python scalar or NumPy array over ranks (comm.allreduce / Allreduce)."""
        arr = np.asarray(x)
        t = torch.as_tensor(arr.astype(np.float64 if arr.dtype.kind == 'f' else np.int64)).to(self._host_device())
        if t.dim() == 0:
            t = t.reshape(1)
            dist.all_reduce(t, group=self.group)
            v = t.cpu().numpy()[0]
            return float(v) if arr.dtype.kind == 'f' else int(v)
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def allreduce_max(self, x):
        t = torch.as_tensor([float(x)], dtype=torch.float64).to(self._host_device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.cpu()[0])

    def bcast(self, x, root=0):
        obj = [x]
        dist.broadcast_object_list(obj, src=root, group=self.group)
        return obj[0]

    def allreduce_tensor_(self, t):
        """In-place sum of a tensor that already lives where the backend wants it."""
        dist.all_reduce(t, group=self.group)
        return t

    def allgather_tensor(self, t):
        """Gather equally-shaped tensors from all ranks."""
        out = [torch.empty_like(t) for _ in range(self.size)]
        dist.all_gather(out, t, group=self.group)
        return out

    def Barrier(self):
        dist.barrier(group=self.group)


def default_comm():
    return TorchComm() if dist.is_available() and dist.is_initialized() else SerialComm()


def init_from_env():
    """Under torchrun (RANK / WORLD_SIZE / MASTER_* in the environment) bring up the process group this package's
    communicator wraps: NCCL with one GPU per rank (cuda:LOCAL_RANK), gloo on a host without GPUs.  The counterpart of
    `mpirun` creating MPI.COMM_WORLD for the reference.  Returns the communicator."""
    import os
    if dist.is_available() and not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if torch.cuda.is_available():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    return default_comm()


def pprint(obj="", comm=None, end='\n'):
    comm = comm or default_comm()
    if comm.rank != 0:
        return
    sys.stdout.write((obj if isinstance(obj, str) else repr(obj)) + end)
    sys.stdout.flush()


def stride_data(N, balanced=False, comm=None):
    """Block distribution of N items over comm.size ranks -> (first, last).

    Same rule as parallel.py:67-84: N // size each, the first N % size ranks get one more
    (`balanced=True` drops the remainder).
    """
    comm = comm or default_comm()
    base, residue = divmod(N, comm.size)
    if balanced:
        return base * comm.rank, base * (comm.rank + 1)
    if comm.rank < residue:
        first = (base + 1) * comm.rank
        return first, first + base + 1
    first = base * comm.rank + residue
    return first, first + base


def _allgather_flat(my_array, axis, comm):
    """Rank-ordered concatenation of equally shaped local arrays as ONE flat buffer, viewed with the length of `axis`
    multiplied by the number of ranks -- what MPI's Allgather into an array of that shape does (parallel.py:104-107)."""
    all_shape = list(my_array.shape)
    all_shape[axis] = int(comm.allreduce(my_array.shape[axis]))
    if comm.size == 1:
        return my_array.reshape(all_shape)
    t = torch.as_tensor(np.ascontiguousarray(my_array)).to(comm._host_device())
    flat = np.concatenate([p.cpu().numpy().ravel() for p in comm.allgather_tensor(t)])
    return flat.reshape(all_shape)


def _check_sortable(my_array):
    if my_array.dtype.kind not in 'fiub':
        raise TypeError("Dont know how to handle arrays of type %s" % my_array.dtype)


def allsort(my_array, axis=-1, kind='quicksort', order=None, comm=None):
    """All ranks get the globally sorted array (parallel.py:87-110): local sort, all-gather, merge sort.

    1-D arrays (the truncation cut of the M-step, bsc_et.py:252) may have a different length on every rank -- the
    reference's Allgather needs equal lengths, `stride_data` shards differ by one.  N-D arrays follow the reference
    exactly: equal shapes on all ranks, the gathered buffer viewed with `axis` lengthened, sorted along `axis`."""
    comm = comm or default_comm()
    my_array = np.asarray(my_array)
    _check_sortable(my_array)
    if my_array.ndim != 1:
        return np.sort(_allgather_flat(np.sort(my_array, axis, kind, order), axis, comm), axis, 'mergesort', order)
    if comm.size == 1:
        return np.sort(my_array)
    n = int(comm.allreduce(len(my_array)))
    nmax = int(comm.allreduce_max(len(my_array)))
    pad = np.full(nmax, np.inf)                 # +inf padding sorts last and is cut off below
    pad[:len(my_array)] = my_array
    t = torch.as_tensor(pad).to(comm._host_device())
    parts = comm.allgather_tensor(t)
    allv = np.sort(np.concatenate([p.cpu().numpy() for p in parts]))
    return allv[:n].astype(my_array.dtype, copy=False)


def allargsort(my_array, axis=-1, kind='quicksort', order=None, comm=None):
    """parallel.py:113-135, as upstream: the argsort of the gathered LOCAL argsort indices (not a global argsort of the
    values -- the reference gathers `np.argsort(my_array)` and argsorts that)."""
    comm = comm or default_comm()
    my_array = np.asarray(my_array)
    _check_sortable(my_array)
    return np.argsort(_allgather_flat(np.argsort(my_array, axis, kind, order), axis, comm), axis, kind, order)


def allmean(my_a, axis=None, dtype=None, out=None, comm=None):
    """parallel.py:138-156."""
    comm = comm or default_comm()
    N = comm.allreduce(my_a.size if axis is None else my_a.shape[axis])
    return comm.allreduce(np.sum(my_a, axis, dtype)) / N


def allsum(my_a, axis=None, dtype=None, out=None, comm=None):
    """parallel.py:159-170."""
    comm = comm or default_comm()
    return comm.allreduce(np.sum(my_a, axis, dtype))


def bind_to_gpu_numa_node(device_index):
    """Restrict the calling thread to the CPUs NVML reports as local to GPU `device_index`.

    One process per GPU: pinned host buffers are first-touched by this thread, so they land on the NUMA node the GPU's
    PCIe link hangs off and host->device copies of several ranks do not cross the socket interconnect (the reference's
    counterpart is `mpirun --bind-to ...`; prosper itself leaves placement to MPI).  Returns the CPU list, or None when
    nothing was changed (no NVML, a single node, a cpuset that excludes the local CPUs)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        handle = None
        uuid = getattr(props, 'uuid', None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-%s" % uuid).encode())
            except Exception:
                handle = None
        if handle is None:
            if os.environ.get("CUDA_VISIBLE_DEVICES"):
                return None                              # indices are remapped and the UUID lookup failed: leave it
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        local = set(64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1)
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if len(target) < 2 or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return sorted(target)
    except Exception:
        return None
