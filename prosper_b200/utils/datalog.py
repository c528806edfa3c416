"""Rank-0 publish/subscribe of named values (the keys the hot path feeds).

Mirrors the interface of prosper/utils/datalog.py (`dlog.set_handler / append / append_all /
progress / close`, datalog.py:179-254) so that the models log the same keys the reference
does (`W, pi, sigma, mu, L, N, N_use, Q, prior_mass` and the annealing values,
camodels/__init__.py:190-191).  Handlers here: TextPrinter, StoreToTxt, StoreToH5 (`result.h5`
through utils/autotable.py + utils/h5min.py), Keep (in-memory).
"""
import sys

import numpy as np


class DataHandler(object):
    def register(self, tblname):
        pass

    def append(self, tblname, value):
        raise NotImplementedError

    def append_all(self, valdict):
        for k, v in valdict.items():
            self.append(k, v)

    def remove(self, tblname):
        pass

    def close(self):
        pass


class TextPrinter(DataHandler):
    def append(self, tblname, value):
        sys.stdout.write("  %8s = %s \n" % (tblname, value))


class StoreToTxt(DataHandler):
    def __init__(self, destination='terminal.txt'):
        self.fd = open(destination, 'w')

    def append(self, tblname, value):
        self.fd.write("%s = %s\n" % (tblname, value))
        self.fd.flush()

    def close(self):
        self.fd.close()


class StoreToH5(DataHandler):
    """Store every appended value as a new row of the table of that name in an HDF5 file
    (datalog.py:53-93): `dlog.set_handler(('W', 'pi', 'sigma', 'L'), StoreToH5, 'output/result.h5')`.
    `destination` is a file name, an AutoTable, or None (shared default table)."""
    default_autotbl = None

    def __init__(self, destination=None):
        from .autotable import AutoTable
        self.destination = destination
        if isinstance(destination, AutoTable):
            self.autotbl = destination
        elif isinstance(destination, str):
            self.autotbl = AutoTable(destination)
        elif destination is None:
            self.autotbl = AutoTable() if StoreToH5.default_autotbl is None else StoreToH5.default_autotbl
        else:
            raise TypeError("Expects an AutoTable instance or a string as argument")
        if StoreToH5.default_autotbl is None:
            StoreToH5.default_autotbl = self.autotbl

    def __repr__(self):
        return "StoreToH5 into file %s" % self.destination

    def append(self, tblname, value):
        self.autotbl.append(tblname, value)

    def append_all(self, valdict):
        self.autotbl.append_all(valdict)

    def close(self):
        self.autotbl.close()


class Keep(DataHandler):
    """Collects every appended value in `self.values[name]` (used by tests and bench)."""

    def __init__(self):
        self.values = {}

    def append(self, tblname, value):
        self.values.setdefault(tblname, []).append(np.copy(value) if isinstance(value, np.ndarray) else value)

    def last(self, name):
        return self.values[name][-1]


class DataLog(object):
    def __init__(self, comm=None):
        self.comm = comm
        self.policy = []          # ordered (tblname or '*', handler)

    def _rank(self):
        if self.comm is not None:
            return self.comm.rank
        from . import parallel
        return parallel.default_comm().rank

    def _lookup(self, tblname):
        return [h for (name, h) in self.policy if name == tblname or name == '*']

    def set_handler(self, tblname, handler_class, *args, **kargs):
        """datalog.py:234-254: one handler instance for the table name (or every name of an iterable)."""
        if self._rank() != 0:
            return None
        if not issubclass(handler_class, DataHandler):
            raise TypeError("handler_class must be a subclass of DataHandler")
        handler = handler_class(*args, **kargs)
        if isinstance(tblname, str):
            tblnames = (tblname,)
        elif hasattr(tblname, '__iter__'):
            tblnames = tuple(tblname)
        else:
            raise TypeError('Table-name must be a string (or a list of strings)')
        for t in tblnames:
            self.policy.append((t, handler))
            handler.register(t)
        return handler

    def remove_handler(self, handler):
        self.policy = [(n, h) for (n, h) in self.policy if h is not handler]

    def ignored(self, tblname):
        return self._rank() != 0 or not self._lookup(tblname)

    def append(self, tblname, value):
        if self._rank() != 0:
            return
        handlers = self._lookup(tblname)
        if not handlers:
            return                       # nobody listens: a device-resident value is never copied to the host
        if hasattr(value, 'detach') and hasattr(value, 'cpu'):
            value = value.detach().cpu().numpy()
        for h in handlers:
            h.append(tblname, value)

    def append_all(self, valdict):
        for k, v in valdict.items():
            self.append(k, v)

    def progress(self, message, completed=None):
        if self._rank() != 0:
            return
        if completed is None:
            sys.stdout.write("[%s]\n" % message)
        else:
            sys.stdout.write("[%5.1f%%] %s\n" % (completed * 100., message))
        sys.stdout.flush()

    def close(self, quiet=False):
        if self._rank() != 0:
            return
        seen = []
        for _, h in self.policy:
            if not any(h is s for s in seen):
                h.close()
                seen.append(h)
        self.policy = []


dlog = DataLog()
