"""Host-side helpers mirroring prosper.utils for the hot path (parallel, datalog, autotable, barstest)."""
import errno
import os
import sys
import time


def create_output_path(basename=None):
    """A fresh output directory "output/<basename>.<suffix>/" made by rank 0 and broadcast (prosper/utils/__init__.py:17-62).

    <suffix> is "d<job id>" under PBS / SLURM, the date and time otherwise; "+<n>" is appended while the name is taken.
    `basename` defaults to the program's name (sys.argv[0])."""
    from . import parallel
    comm = parallel.default_comm()
    dirname = None
    if comm.rank == 0:
        if basename is None:
            basename = sys.argv[0]
        if 'PBS_JOBID' in os.environ:
            suffix = "d" + os.environ['PBS_JOBID'].split('.')[0]
        elif 'SLURM_JOBID' in os.environ:
            suffix = "d" + os.environ['SLURM_JOBID']
        else:
            suffix = time.strftime("%Y-%m-%d+%H:%M")
        dirname, tries = "output/%s.%s" % (basename, suffix), 0
        while True:
            try:
                os.makedirs(dirname)
                break
            except OSError as e:
                if e.errno != errno.EEXIST:
                    raise
                tries += 1
                dirname = "output/%s.%s+%d" % (basename, suffix, tries)
    return comm.bcast(dirname) + "/"
