"""Import-time stand-in for PyTables (TEST INFRASTRUCTURE ONLY); see ../mpi4py."""


class Filters(object):
    def __init__(self, *a, **k):
        pass


def open_file(*a, **k):
    raise RuntimeError("PyTables is not available in this image (shim)")


openFile = open_file
