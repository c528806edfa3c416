"""result.h5 writer: file structure, AutoTable row semantics, the dlog handler."""
import os
import struct

import numpy as np
import pytest

from prosper_b200.utils import h5min
from prosper_b200.utils.autotable import AutoTable
from prosper_b200.utils.datalog import DataLog, StoreToH5, Keep


def test_roundtrip_all_supported_types(tmp_path):
    rng = np.random.RandomState(0)
    d = {'W': rng.randn(3, 4, 5), 'pi': np.arange(3.0), 'N_use': np.array([5, 6, 7]), 'flag': np.array([True, False]),
         'name': np.array([b'abc', b'de']), 'x32': np.arange(6, dtype=np.float32).reshape(2, 3), 'empty': np.zeros((0, 4)),
         'i8': np.array([[-1, 0, 1]], dtype=np.int8), 'u16': np.array([1, 65535], dtype=np.uint16), 'scalar': np.array(3.5)}
    path = str(tmp_path / "t.h5")
    h5min.write_h5(path, d)
    r = h5min.read_h5(path)
    assert sorted(r) == sorted(d)
    for k in d:
        a = np.asarray(d[k])
        assert r[k].shape == a.shape and (r[k] == a.astype(r[k].dtype)).all(), k
    assert r['W'].dtype == np.float64 and r['N_use'].dtype == np.int64 and r['x32'].dtype == np.float32


def test_file_structure_follows_the_hdf5_specification(tmp_path):
    """Spot checks of the on-disk bytes against the HDF5 file format specification (v0 superblock, v1 objects)."""
    path = str(tmp_path / "s.h5")
    h5min.write_h5(path, {'b': np.arange(4.0), 'a': np.arange(6, dtype=np.int64).reshape(2, 3)})
    raw = open(path, 'rb').read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert raw[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])                 # versions, 8-byte offsets and lengths
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert (base, free, eof, drv) == (0, h5min.UNDEF, len(raw), h5min.UNDEF)
    root_hdr = struct.unpack_from("<Q", raw, 64)[0]
    assert struct.unpack_from("<I", raw, 72)[0] == 1                    # cache type 1: scratch pad holds B-tree / heap
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    assert raw[root_hdr] == 1 and struct.unpack_from("<H", raw, root_hdr + 16)[0] == 0x0011   # symbol table message
    assert struct.unpack_from("<Q", raw, heap + 16)[0] == 1             # libhdf5's "no free block" marker
    snod = struct.unpack_from("<Q", raw, btree + 32)[0]
    assert raw[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", raw, snod + 6)[0] == 2
    # IEEE float64 little endian datatype message exactly as libhdf5 writes H5T_IEEE_F64LE
    assert bytes.fromhex("11203f0008000000000040003 40b0034ff030000".replace(" ", "")) in raw
    # signed 64-bit little endian integer (H5T_STD_I64LE)
    assert bytes.fromhex("100800000800000000004000") in raw
    r = h5min.read_h5(path)
    assert list(r) == ['a', 'b']                                        # symbol table sorted by name


def test_many_datasets_fit_one_symbol_table_node(tmp_path):
    d = dict(('k%02d' % i, np.full((2, i + 1), float(i))) for i in range(40))
    path = str(tmp_path / "m.h5")
    h5min.write_h5(path, d)
    r = h5min.read_h5(path)
    assert sorted(r) == sorted(d) and all((r[k] == d[k]).all() for k in d)


def test_autotable_appends_rows_like_the_reference(tmp_path):
    """autotable.py:87-127: table shape = (number of appends,) + value.shape; scalars become 1-d tables."""
    path = str(tmp_path / "result.h5")
    with AutoTable(path) as tbl:
        for it in range(5):
            tbl.append('W', np.full((3, 2), float(it)))
            tbl.append('pi', 0.1 * it)
            tbl.append('N_use', 10 + it)
        tbl.append_all({'L': -3.5, 'sigma': 1.0})
        with pytest.raises(TypeError):
            tbl.append('W', np.zeros((2, 2)))
        with pytest.raises(TypeError):
            tbl.append('W', object())
    r = h5min.read_h5(path)
    assert r['W'].shape == (5, 3, 2) and (r['W'][3] == 3.0).all()
    assert r['pi'].shape == (5,) and np.allclose(r['pi'], 0.1 * np.arange(5))
    assert r['N_use'].dtype.kind == 'i' and list(r['N_use']) == [10, 11, 12, 13, 14]
    assert r['L'].shape == (1,) and r['sigma'][0] == 1.0


def test_dlog_store_to_h5(tmp_path):
    """datalog.py:53-93,179-254: the handler receives only the tables it was registered for."""
    path = str(tmp_path / "out" / "result.h5")
    os.makedirs(os.path.dirname(path))
    log = DataLog()
    log.set_handler(('W', 'pi', 'L'), StoreToH5, path)
    keep = log.set_handler('*', Keep)
    for it in range(3):
        log.append_all({'W': np.eye(2) * it, 'pi': 0.2, 'sigma': 1.5})
        log.append('L', -1.0 * it)
    assert not log.ignored('W') and not log.ignored('sigma')
    log.close()
    r = h5min.read_h5(path)
    assert sorted(r) == ['L', 'W', 'pi']
    assert r['W'].shape == (3, 2, 2) and r['L'].tolist() == [0.0, -1.0, -2.0]
    assert len(keep.values['sigma']) == 3


LIBHDF5_FILE = os.path.join(os.path.dirname(np.__file__), "..", "scipy", "io", "matlab", "tests", "data",
                            "testhdf5_7.4_GLNX86.mat")


def _object_messages(raw, addr):
    """(type, body) of the v1 object header at absolute address addr (first block only)."""
    ver, _, nmsg, _, size = struct.unpack_from("<BBHII", raw, addr)
    assert ver == 1
    p, out = addr + 16, []
    while p < addr + 16 + size and len(out) < nmsg:
        mtype, msize, _flags = struct.unpack_from("<HHB", raw, p)
        out.append((mtype, raw[p + 8:p + 8 + msize]))
        p += 8 + msize
    return out


@pytest.mark.skipif(not os.path.exists(LIBHDF5_FILE), reason="SciPy's test data (a file written by libhdf5) is not installed")
def test_reader_and_writer_against_a_file_written_by_libhdf5(tmp_path):
    """Neither libhdf5 nor PyTables is in the image, but SciPy ships a MATLAB 7.4 v7.3 file, i.e. genuine libhdf5 1.6
    output (512-byte user block, v0 superblock, symbol-table root group, one contiguous float64 dataset 0:pi/4:2pi).
    (a) the reader that checks our writer parses it, (b) the messages our writer emits for the same array are byte-identical
    to libhdf5's where the format leaves no choice (superblock versions and sizes, datatype, dataspace)."""
    want = np.linspace(0.0, 2 * np.pi, 9).reshape(9, 1)
    got = h5min.read_h5(LIBHDF5_FILE)
    assert list(got) == ['testdouble'] and got['testdouble'].dtype == np.float64
    assert np.allclose(got['testdouble'], want, rtol=0, atol=1e-15)

    ours_path = str(tmp_path / "ours.h5")
    h5min.write_h5(ours_path, {'testdouble': got['testdouble']})
    ours, ref = open(ours_path, 'rb').read(), open(LIBHDF5_FILE, 'rb').read()
    assert ours[:16] == ref[512:528]                           # signature, all version bytes, offset / length sizes
    # libhdf5 stores the same symbol-table leaf K (4) and B-tree internal K (16) that this writer declares
    assert struct.unpack_from("<HH", ours, 16) == struct.unpack_from("<HH", ref, 512 + 16)

    def dataset_messages(raw, base):
        root_hdr = struct.unpack_from("<Q", raw, base + 64)[0] + base
        btree = struct.unpack_from("<Q", raw, base + 80)[0] + base
        assert dict(_object_messages(raw, root_hdr))[0x0011][:8] == struct.pack("<Q", btree - base)
        snod = struct.unpack_from("<Q", raw, btree + 32)[0] + base
        assert raw[btree:btree + 4] == b"TREE" and raw[snod:snod + 4] == b"SNOD"
        ohdr = struct.unpack_from("<Q", raw, snod + 8 + 8)[0] + base
        return dict(_object_messages(raw, ohdr))

    mo, mr = dataset_messages(ours, 0), dataset_messages(ref, 512)
    assert mo[0x0003] == mr[0x0003]                            # H5T_IEEE_F64LE, byte for byte
    assert mo[0x0001] == mr[0x0001]                            # dataspace: version 1, rank 2, dims (9, 1)
    assert mo[0x0008][0] == 3 and mr[0x0008][0] == 2           # layout: we write version 3 (libhdf5 >= 1.6.3), MATLAB's 1.6 wrote 2
    assert (h5min.read_h5(ours_path)['testdouble'] == got['testdouble']).all()


def test_appendlist_adds_one_row_per_entry(tmp_path):
    """autotable.py:190-223: appendList(name, array) appends array.shape[0] rows; a list of strings one row each."""
    path = str(tmp_path / "l.h5")
    with AutoTable(path) as tbl:
        tbl.append('x', np.zeros(3))
        tbl.appendList('x', np.arange(6.0).reshape(2, 3))
        tbl.appendList('names', ['ab', 'c'])
        tbl.appendList('t', np.arange(4.0))
        with pytest.raises(TypeError):
            tbl.appendList('x', [1, 2, 3])
    r = h5min.read_h5(path)
    assert r['x'].shape == (3, 3) and (r['x'][1:] == np.arange(6.0).reshape(2, 3)).all()
    assert list(r['names']) == [b'ab', b'c'] and (r['t'] == np.arange(4.0)).all()


def test_set_handler_accepts_a_name_or_an_iterable_of_names():
    """datalog.py:234-254."""
    log = DataLog()
    h = log.set_handler(('a', 'b'), Keep)
    log.append('a', 1.0); log.append('b', 2.0); log.append('c', 3.0)
    assert h.values['a'] == [1.0] and h.values['b'] == [2.0] and 'c' not in h.values
    with pytest.raises(TypeError):
        log.set_handler(5, Keep)
    with pytest.raises(TypeError):
        log.set_handler('a', dict)


def test_reference_example_flow_with_wildcard_handlers(tmp_path, capsys):
    """The logging flow of the reference's examples/barstests/bars-learning.py (TextPrinter + StoreToTxt on a print list,
    StoreToH5 on '*', EM.run over a LinearAnnealing schedule, dlog.close) with a model whose step() logs what
    CAModel.step logs (camodels/__init__.py:190-191) -- everything but the CUDA step, so it runs without a GPU."""
    from prosper_b200.em import EM, Model
    from prosper_b200.em.annealing import LinearAnnealing
    from prosper_b200.utils.datalog import dlog, TextPrinter, StoreToTxt

    class Stub(Model):
        def step(self, anneal, model_params, my_data):
            new = {'W': model_params['W'] + 1.0, 'pi': float(model_params['pi']) * 0.5, 'sigma': 1.0,
                   'mu': np.zeros(4), 'Q': 0.}
            dlog.append('L', -1.5)
            dlog.append('N', 10)
            dlog.append_all(new)
            dlog.append_all(anneal.as_dict())
            return new

    anneal = LinearAnnealing(5)
    anneal['T'] = [(0, 2.), (.7, 1.)]
    anneal['Ncut_factor'] = [(0, 0.), (2. / 3, 1.)]
    anneal['anneal_prior'] = False
    print_list = ('T', 'Q', 'pi', 'sigma', 'N', 'MAE')
    handlers = [dlog.set_handler(print_list, TextPrinter),
                dlog.set_handler(print_list, StoreToTxt, str(tmp_path / "terminal.txt")),
                dlog.set_handler(('*'), StoreToH5, str(tmp_path / "result.h5"))]
    try:
        em = EM(model=Stub(), anneal=anneal)
        em.data = {'y': np.zeros((10, 4))}
        em.lparams = {'W': np.zeros((4, 3)), 'pi': 0.5, 'sigma': 1.0}
        em.run()
        dlog.close()
    finally:
        for h in handlers:
            dlog.remove_handler(h)
    r = h5min.read_h5(str(tmp_path / "result.h5"))
    assert r['W'].shape == (5, 4, 3) and (r['W'][:, 0, 0] == np.arange(1.0, 6.0)).all()
    assert r['pi'].shape == (5,) and r['T'][0] == 2.0 and r['T'][-1] == 1.0 and r['N'].tolist() == [10] * 5
    assert r['Ncut_factor'][0] == 0.0 and r['Ncut_factor'][-1] == 1.0 and 'anneal_prior' in r and 'L' in r
    assert (em.lparams['W'] == 5.0).all()
    assert "sigma" in open(str(tmp_path / "terminal.txt")).read() and "pi" in capsys.readouterr().out


def test_tracepoints_of_the_stage_methods(tmp_path):
    """utils/tracing.py (tracing.py:38-141): set_tracefile / traced / close, one file per rank packed into traces.tgz."""
    import tarfile
    from prosper_b200.utils import tracing
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    assert BSC_ET.E_step.__name__ == 'E_step' and BSC_ET.step.__wrapped__ is not None

    @tracing.traced
    def work(x):
        """doc"""
        return x + 1

    assert work(1) == 2                                   # inactive: plain call
    tracing.set_tracefile(str(tmp_path / "trace-%04d.txt"))
    assert work(2) == 3
    tracing.tracepoint("custom")
    tracing.close()
    assert tracing.trace_file is None and not (tmp_path / "trace-0000.txt").exists()
    with tarfile.open(str(tmp_path / "traces.tgz")) as tar:
        text = tar.extractfile("trace-0000.txt").read().decode()
    assert "[work:begin]" in text and "[work:end]" in text and "[custom]" in text and text.startswith("# Start time")


def test_autotable_flushes_while_the_run_is_going(tmp_path):
    """The reference appends to on-disk EArrays (autotable.py:87-127), so a killed run keeps its log.  Here rows are
    buffered and the file is rewritten atomically every `flush_interval` seconds from inside `append`."""
    from prosper_b200.utils.autotable import AutoTable
    from prosper_b200.utils import h5min
    fn = str(tmp_path / "run.h5")
    t = AutoTable(fn, flush_interval=1e-9)                 # every append is "due"
    t.append('pi', 0.25)
    t.append('W', np.ones((3, 2)))
    got = h5min.read_h5(fn)                                # no close(), no flush(): the rows are on disk already
    assert got['pi'].shape == (1,) and got['W'].shape == (1, 3, 2)
    t.append('pi', 0.5)
    assert h5min.read_h5(fn)['pi'].tolist() == [0.25, 0.5]
    t2 = AutoTable(str(tmp_path / "lazy.h5"), flush_interval=None)
    t2.append('pi', 0.25)
    assert not (tmp_path / "lazy.h5").exists()
    t2.close()
    assert h5min.read_h5(str(tmp_path / "lazy.h5"))['pi'].tolist() == [0.25]
