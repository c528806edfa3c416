"""prosper_b200 -- B200-native truncated-EM (Expectation Truncation) engine behind prosper's API.

Mirrors the reference's operator interface for ONE hot path (select_Hprimes -> E_step ->
M_step of prosper/em/camodels/*_et.py); see DESIGN.md.  The arithmetic runs in hand-written
sm_100a CUDA kernels reached through the C ABI of include/prosper_b200.h; there is no CPU
fallback -- constructing a model without the built library or without a B200 raises.
"""
import importlib.util

__version__ = "0.1.0"

_PROSPER_MODULES = ("em", "em.annealing", "em.camodels", "em.camodels.bsc_et", "em.camodels.mca_et", "em.camodels.mmca_et",
                    "em.camodels.tsc_et", "em.camodels.dsc_et", "em.camodels.gsc_et", "em.mixturemodels",
                    "em.mixturemodels.MoG", "em.mixturemodels.MoP", "utils", "utils.parallel", "utils.datalog",
                    "utils.autotable", "utils.barstest", "utils.tracing")


def install_as_prosper(mpi_shim=True):
    """Register this package under the name `prosper`, so that scripts written for the reference
    (`from prosper.em.camodels.bsc_et import BSC_ET`, `from prosper.utils.datalog import dlog`, ... -- every import of
    the reference's examples/) run on this engine unchanged:

        import prosper_b200; prosper_b200.install_as_prosper()

    The reference's scripts also do `from mpi4py import MPI` for `MPI.COMM_WORLD.rank / .size`.  With `mpi_shim` and no
    real mpi4py installed, a stand-in module is registered whose COMM_WORLD is this package's communicator
    (torch.distributed under torchrun -- brought up here from the environment -- or the single-process one).
    Refuses to shadow a real `prosper` that is already imported."""
    import importlib
    import sys
    me = sys.modules[__name__]
    other = sys.modules.get("prosper")
    if other is not None and other is not me:
        raise RuntimeError("a different 'prosper' package is already imported")
    sys.modules["prosper"] = me
    for name in _PROSPER_MODULES:
        sys.modules["prosper." + name] = importlib.import_module(__name__ + "." + name)
    if mpi_shim and "mpi4py" not in sys.modules and importlib.util.find_spec("mpi4py") is None:
        import time
        import types
        from .utils import parallel
        parallel.init_from_env()
        pkg, mpi = types.ModuleType("mpi4py"), types.ModuleType("mpi4py.MPI")
        mpi.__getattr__ = lambda name: parallel.default_comm() if name == "COMM_WORLD" else _no_mpi(name)
        mpi.Wtime = time.time
        pkg.MPI = mpi
        pkg.__prosper_b200_shim__ = True
        sys.modules["mpi4py"], sys.modules["mpi4py.MPI"] = pkg, mpi
    return me


def _no_mpi(name):
    raise AttributeError("mpi4py stand-in of prosper_b200 has no attribute %r (only COMM_WORLD and Wtime)" % name)
