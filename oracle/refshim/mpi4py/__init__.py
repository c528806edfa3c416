"""Single-rank stand-in for mpi4py (TEST INFRASTRUCTURE ONLY).

Lets the unmodified reference under /root/reference be imported in the build
container, where neither mpi4py nor an MPI runtime exist, so that golden
vectors can be minted from it (tests/golden/make_golden.py).  Same idea as the
reference's own Sphinx mocks (docs/conf.py:19-27).  Never imported by the
product package.
"""
import time as _time
import numpy as _np


class _Comm(object):
    rank = 0
    size = 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    # pickle-style collectives: identity on one rank
    def allreduce(self, x, op=None):
        return x

    def bcast(self, x, root=0):
        return x

    def allgather(self, x):
        return [x]

    def gather(self, x, root=0):
        return [x]

    # buffer-style collectives: copy send -> recv
    @staticmethod
    def _buf(spec):
        return spec[0] if isinstance(spec, (tuple, list)) else spec

    def Allreduce(self, send, recv, op=None):
        _np.copyto(self._buf(recv), self._buf(send))

    def Allgather(self, send, recv):
        r = self._buf(recv)
        r[...] = _np.asarray(self._buf(send)).reshape(r.shape)

    def Bcast(self, buf, root=0):
        pass

    def Barrier(self):
        pass


class MPI(object):
    COMM_WORLD = _Comm()
    DOUBLE = 'DOUBLE'
    FLOAT = 'FLOAT'
    SHORT = 'SHORT'
    INT = 'INT'
    LONG = 'LONG'
    UNSIGNED_SHORT = 'UNSIGNED_SHORT'
    UNSIGNED_INT = 'UNSIGNED_INT'
    UNSIGNED_LONG = 'UNSIGNED_LONG'
    SUM = 'SUM'
    Wtime = staticmethod(_time.time)
