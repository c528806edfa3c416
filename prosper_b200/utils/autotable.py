"""AutoTable: append-only named tables stored in one HDF5 file (prosper/utils/autotable.py:40-278).

Same interface as the reference for the logging path (`append`, `append_all`, `assign`, `close`,
context manager): every `append(name, value)` adds one row, so after T iterations `/W` has shape
(T, D, H), `/pi` (T,), ... -- the layout the reference's notebooks read with
`tables.open_file('result.h5').root.W[:]`.  PyTables is not available here; rows are kept in host
memory and the file is written by `utils/h5min.py` on `flush()` / `close()` (flat root group,
contiguous datasets) and, so that a run that is killed keeps what it has logged (the reference appends to
on-disk EArrays as it goes), every `flush_interval` seconds from inside `append` (atomic: `.tmp` + `os.replace`).
"""
import os
import time

import numpy as np

from . import h5min


class AutoTable(object):
    def __init__(self, fname=None, compression_level=1, flush_interval=60.0):
        if fname is None:
            fname = self._guess_fname()
        self.fname = fname
        self.flush_interval = flush_interval             # seconds between automatic flushes; None / <= 0: only on close()
        self._last_flush = time.time()
        self.compression_level = compression_level       # accepted for compatibility; datasets are stored uncompressed
        self.tables = {}
        self.types = {}
        self._closed = False

    def __enter__(self):
        return self

    def __exit__(self, *exc_info):
        self.close()

    @staticmethod
    def _guess_fname():
        """autotable.py:280-295: <script name>.<timestamp>.h5 style default."""
        import sys
        base = os.path.splitext(os.path.basename(sys.argv[0] or "autotable"))[0] or "autotable"
        return "%s.%s.h5" % (base, time.strftime("%Y-%m-%d+%H:%M"))

    def append(self, name, value):
        """autotable.py:87-127 (same type rules and errors)."""
        if type(value) == np.ma.core.MaskedArray:
            value = value.data
        if type(value) == str:
            value = np.asarray(value.encode("ascii", "replace"))
        elif np.isscalar(value):
            value = np.asarray(value)
        if not isinstance(value, np.ndarray):
            raise TypeError("Don't know how to handle values of type '%s'", type(value))
        if name not in self.tables:
            self.tables[name] = []
            self.types[name] = (value.dtype, value.shape)
        dt, shape = self.types[name]
        if value.shape != shape or (value.dtype.kind != dt.kind and not (dt.kind in 'fiu' and value.dtype.kind in 'fiub')):
            raise TypeError('Wrong datatype "%s" for "%s" field' % (value.dtype, name))
        self.tables[name].append(np.array(value, copy=True))
        if self.flush_interval and self.flush_interval > 0 and time.time() - self._last_flush >= self.flush_interval:
            self.flush()

    def appendList(self, name, value):
        """autotable.py:190-223: like `append`, but `value` holds SEVERAL rows (its first axis, or a list of strings)."""
        if type(value) == list and len(value) > 0 and type(value[0]) == str:
            for v in value:
                self.append(name, v)
            return
        if np.isscalar(value):
            value = np.asarray(value)
        if not isinstance(value, np.ndarray):
            raise TypeError("Don't know how to handle values of type '%s'", type(value))
        for row in (value if value.ndim > 0 else value.reshape(1)):
            self.append(name, row)

    def append_all(self, valdict):
        """autotable.py:156-166."""
        for name, value in valdict.items():
            self.append(name, value)

    def assign(self, name, value):
        """autotable.py:129-154: replace the whole table by `value` (first axis = rows)."""
        value = np.asarray(value)
        self.tables[name] = [np.array(v, copy=True) for v in value]
        self.types[name] = (value.dtype, value.shape[1:])

    def _stacked(self):
        out = {}
        for name, rows in self.tables.items():
            dt, shape = self.types[name]
            if dt.kind in 'SU':
                width = max([1] + [r.dtype.itemsize for r in rows])
                out[name] = np.array([r.item() for r in rows], dtype='S%d' % width)
            elif rows:
                out[name] = np.stack(rows).astype(dt if dt.kind != 'b' else np.uint8, copy=False)
            else:
                out[name] = np.zeros((0,) + shape, dtype=dt)
        return out

    def flush(self):
        h5min.write_h5(self.fname + ".tmp", self._stacked())
        os.replace(self.fname + ".tmp", self.fname)
        self._last_flush = time.time()

    def close(self):
        if not self._closed:
            self.flush()
            self._closed = True
