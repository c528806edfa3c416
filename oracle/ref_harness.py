"""Import the UNMODIFIED reference (read-only /root/reference) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py to mint golden
vectors and by the optional `-m "not gpu"` live-reference tests; skipped
wherever /root/reference does not exist (e.g. the GPU box).  Recipe follows
SURVEY.md Appendix C: fake mpi4py + tables on sys.path, NumPy-2 aliases
(np.int & co. are used at bsc_et.py:109, mca_et.py:72, ...), and the `states`
global that tsc_et.py:131 looks up.
"""
import os
import sys
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PROSPER_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "prosper"))


def load():
    """Return the imported reference `prosper` package (raises if absent)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name, val in (("int", int), ("bool", bool), ("str", str), ("object", object), ("float", float)):
        if name not in np.__dict__:
            setattr(np, name, val)
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import prosper  # noqa: F401
    import prosper.em  # noqa: F401
    import prosper.em.annealing  # noqa: F401
    import prosper.em.camodels.bsc_et  # noqa: F401
    import prosper.em.camodels.mca_et  # noqa: F401
    import prosper.em.camodels.mmca_et  # noqa: F401
    import prosper.em.camodels.tsc_et as tsc
    import prosper.em.camodels.dsc_et  # noqa: F401
    import prosper.em.camodels.gsc_et  # noqa: F401
    tsc.states = np.array([-1., 0., 1.])
    return prosper


class KeepLog(object):
    """Collects what the reference models send to `dlog` (L, N, N_use, ...)."""

    def __init__(self):
        self.values = {}

    def install(self):
        from prosper.utils.datalog import dlog, DataHandler
        outer = self

        class _H(DataHandler):
            def append(self, tblname, value):
                outer.values.setdefault(tblname, []).append(value)

        self._handler = dlog.set_handler('*', _H)
        return self

    def last(self, key):
        return self.values[key][-1]
