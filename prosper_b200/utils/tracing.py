"""Tracepoints for run-time profiling (prosper/utils/tracing.py:38-141): same calls, same file format.

    import prosper.utils.tracing as tracing          # after prosper_b200.install_as_prosper()
    tracing.set_tracefile("trace-%04d.txt")          # on every rank
    ... EM ...                                        # select_Hprimes / E_step / M_step / step record :begin / :end
    tracing.close()                                   # rank 0 packs the per-rank files into traces.tgz

Host-side wall-clock marks only; the device-side split of a step (score GEMM, posterior kernels, statistics GEMM, solve)
comes from CUDA events: `Engine.enable_timing()` / `Engine.stage_times()` (`pet_stage_times_ms`)."""
import os
import os.path as path
import tarfile
import time
from functools import wraps

trace_fname = None
trace_file = None
start_time = None


def _comm(comm):
    from . import parallel
    return comm or parallel.default_comm()


def tracepoint(str):
    """Record the tracepoint *str* (a no-op unless `set_tracefile` was called)."""
    if trace_file is None:
        return
    trace_file.write("[%f] [%s]\n" % (time.time() - start_time, str))


def traced(func):
    """Decorator: records "<name>:begin" / "<name>:end" around every call; keeps the function's name and doc."""
    begin_str, end_str = func.__name__ + ':begin', func.__name__ + ':end'

    @wraps(func)
    def wrapped(*args, **kwargs):
        if trace_file is None:
            return func(*args, **kwargs)
        tracepoint(begin_str)
        res = func(*args, **kwargs)
        tracepoint(end_str)
        return res

    return wrapped


def set_tracefile(fname="trace-%04d.txt", comm=None):
    """Enable tracing; `fname` holds a %d that becomes the rank.  Collective: call it on every rank."""
    global trace_fname, trace_file, start_time
    comm = _comm(comm)
    trace_fname = fname
    trace_file = open(fname % comm.rank, "w")
    trace_file.write("# Start time: %s\n" % time.asctime())
    trace_file.write("# Hostname: %s\n" % os.uname()[1])
    trace_file.write("# MPI size: %d rank: %d\n" % (comm.size, comm.rank))
    comm.Barrier()
    start_time = time.time()


def close(archive=True, comm=None):
    """Close the trace files; rank 0 archives them as traces.tgz next to them and removes the originals."""
    global trace_fname, trace_file, start_time
    if trace_file is None:
        return
    comm = _comm(comm)
    tracepoint("closing tracefiles")
    trace_file.close()
    comm.Barrier()
    if archive and comm.rank == 0:
        trace_dir, tail = path.split(trace_fname)
        trace_dir = path.normpath(trace_dir)
        with tarfile.open(path.join(trace_dir, "traces.tgz"), "w:gz") as tar:
            for rank in range(comm.size):
                tar.add(path.join(trace_dir, tail % rank), arcname=tail % rank)
        for rank in range(comm.size):
            os.remove(trace_fname % rank)
    trace_fname = trace_file = start_time = None
