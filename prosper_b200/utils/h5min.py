"""Minimal HDF5 writer (and reader for tests): one flat root group of contiguous datasets.

The reference logs through PyTables (`utils/autotable.py`, `tables.open_file(...)`), which is not
installable here, so `result.h5` is written directly in the HDF5 file format (spec version 1.x objects,
readable by any libhdf5 >= 1.6, hence by PyTables / h5py): superblock v0, the root group as a
symbol-table group (one v1 B-tree node, one symbol-table node, one local heap), every dataset a v1
object header with dataspace, datatype, fill-value and contiguous-layout messages.  PyTables opens such
datasets as `Array`s, so `h5.root.W[:]` of the reference's notebooks works unchanged.

Supported dtypes: float64/float32, (u)int8/16/32/64, bool (stored as uint8), fixed-length byte strings.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
HEAP_FREE_NULL = 1          # libhdf5's on-disk "no free block" marker (H5HL_FREE_NULL)


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _datatype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        size = dt.itemsize
        if size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        elif size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        else:
            raise TypeError("unsupported float size %d" % size)
        # class 1 (floating point), version 1; little endian, implied-msb mantissa normalisation, sign bit position
        return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, size) + props
    if dt.kind in 'iu':
        bits0 = 0x08 if dt.kind == 'i' else 0x00                       # bit 3: two's complement signed
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == 'S':
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)  # class 3 string, null padded, ASCII
    raise TypeError("unsupported dtype %s" % dt)


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _canonical(arr):
    arr = np.asarray(arr)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if arr.dtype.kind == 'U':
        arr = arr.astype('S')
    if arr.dtype.byteorder == '>':
        arr = arr.astype(arr.dtype.newbyteorder('<'))
    return np.ascontiguousarray(arr) if arr.ndim else arr          # (ascontiguousarray would turn 0-d into 1-d)


def write_h5(path, datasets):
    """Write {name: ndarray} as datasets of the root group of a new HDF5 file."""
    names = sorted(datasets.keys(), key=lambda s: s.encode("ascii"))
    for nm in names:
        if "/" in nm or not nm:
            raise ValueError("dataset names must be non-empty and must not contain '/': %r" % nm)
    arrays = [_canonical(datasets[nm]) for nm in names]
    n = len(names)
    leaf_k = max(4, (n + 1) // 2)              # a symbol-table node holds 2 * leaf_k entries: one node is enough
    internal_k = 16

    # ---- local heap data: "" at offset 0, then the names --------------------------------------
    heap_data = b"\x00" * 8
    name_off = []
    for nm in names:
        name_off.append(len(heap_data))
        heap_data += _pad8(nm.encode("ascii") + b"\x00")

    # ---- layout of the file --------------------------------------------------------------------
    pos = 96                                                           # superblock v0
    root_hdr_addr = pos
    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", 0, 0))])      # patched below
    pos += len(root_hdr)
    btree_addr = pos
    btree_size = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
    pos += btree_size
    heap_addr = pos
    pos += 32
    heap_data_addr = pos
    pos += len(heap_data)
    snod_addr = pos
    snod_size = 8 + 2 * leaf_k * 40
    pos += snod_size
    hdr_addr, hdr_bytes, data_addr = [], [], []
    for arr in arrays:
        hdr_addr.append(pos)
        dims = arr.shape
        dataspace = struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)
        msgs = [_message(0x0001, dataspace), _message(0x0003, _datatype_message(arr.dtype), flags=1),
                _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),
                _message(0x0008, struct.pack("<BBQQ", 3, 1, 0, 0))]   # address patched below
        hdr_bytes.append(msgs)
        pos += len(_object_header(msgs))
    for arr in arrays:
        pos += -pos % 8
        data_addr.append(pos if arr.nbytes else UNDEF)
        pos += arr.nbytes
    eof = pos

    with open(path, "wb") as f:
        # superblock
        f.write(SIGNATURE)
        f.write(struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0))
        f.write(struct.pack("<HHI", leaf_k, internal_k, 0))
        f.write(struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF))
        f.write(struct.pack("<QQII", 0, root_hdr_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr))
        assert f.tell() == 96
        # root group object header: symbol table message
        f.write(_object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]))
        # B-tree node (group, leaf level): one child = the symbol-table node
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if n else 0, UNDEF, UNDEF)
        if n:
            node += struct.pack("<QQQ", 0, snod_addr, name_off[-1])
        f.write(node + b"\x00" * (btree_size - len(node)))
        # local heap header + data
        f.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), HEAP_FREE_NULL, heap_data_addr))
        f.write(heap_data)
        # symbol-table node
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, n)
        for i in range(n):
            snod += struct.pack("<QQII16x", name_off[i], hdr_addr[i], 0, 0)
        f.write(snod + b"\x00" * (snod_size - len(snod)))
        # dataset headers
        for i, arr in enumerate(arrays):
            assert f.tell() == hdr_addr[i]
            msgs = hdr_bytes[i]
            msgs[3] = _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr[i], arr.nbytes))
            f.write(_object_header(msgs))
        # raw data
        for i, arr in enumerate(arrays):
            f.write(b"\x00" * (-f.tell() % 8))
            if arr.nbytes:
                assert f.tell() == data_addr[i]
                f.write(arr.tobytes())
        assert f.tell() == eof


# ---- reader (tests; follows the file structure from the superblock, not the writer's bookkeeping) -------
def _read_datatype(buf):
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, 0)
    cls, version = cv & 0x0F, cv >> 4
    assert version == 1
    if cls == 1:
        assert (b0 & 1) == 0, "big-endian floats not supported"
        return np.dtype("<f%d" % size)
    if cls == 0:
        return np.dtype("<%s%d" % ("i" if (b0 & 0x08) else "u", size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise TypeError("datatype class %d" % cls)


def read_h5(path):
    """{name: ndarray} of the root group of a file written by write_h5 -- or by libhdf5 itself, as long as it uses the
    same old-style objects (v0 superblock, symbol-table root group, v1 object headers, contiguous datasets; a user block
    in front of the superblock and the v1/v2 layout message of libhdf5 1.6 are understood).  tests/test_h5_cpu.py reads a
    file written by libhdf5 (MATLAB 7.4) with it, which pins this reader -- and through it the writer -- to the real format."""
    with open(path, "rb") as f:
        raw = f.read()
    sb = 0
    while raw[sb:sb + 8] != SIGNATURE:                       # the superblock may sit behind a user block of 512 * 2^k bytes
        sb = 512 if sb == 0 else 2 * sb
        assert sb < len(raw), "not an HDF5 file"
    sb_ver, _, _, _, _, so, sl, _ = struct.unpack_from("<BBBBBBBB", raw, sb + 8)
    assert sb_ver == 0 and so == 8 and sl == 8
    leaf_k, internal_k, _ = struct.unpack_from("<HHI", raw, sb + 16)
    base, _, eof, _ = struct.unpack_from("<QQQQ", raw, sb + 24)
    assert base == sb and eof in (len(raw), len(raw) - base)
    _, root_hdr, cache_type, _ = struct.unpack_from("<QQII", raw, sb + 56)
    btree_addr, heap_addr = struct.unpack_from("<QQ", raw, sb + 80)
    root_hdr += base
    btree_addr += base
    heap_addr += base

    def messages(addr):
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", raw, addr)
        assert ver == 1
        blocks, out, total = [(addr + 16, size)], [], 0
        while blocks:
            p, left = blocks.pop(0)
            while left > 0 and total < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", raw, p)
                body = raw[p + 8:p + 8 + msize]
                out.append((mtype, body))
                total += 1
                if mtype == 0x0010:                          # continuation block
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr + base, clen))
                p += 8 + msize
                left -= 8 + msize
            assert left >= 0
        assert total == nmsg
        return out

    sym = dict(messages(root_hdr))[0x0011]
    assert struct.unpack("<QQ", sym[:16]) == (btree_addr - base, heap_addr - base)
    assert raw[heap_addr:heap_addr + 4] == b"HEAP"
    heap_size, free_head, heap_data = struct.unpack_from("<QQQ", raw, heap_addr + 8)
    heap_data += base
    assert free_head == HEAP_FREE_NULL or free_head < heap_size

    def name_at(off):
        end = raw.index(b"\x00", heap_data + off)
        return raw[heap_data + off:end].decode("ascii")

    out = {}

    def walk(addr):
        assert raw[addr:addr + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", raw, addr + 4)
        assert ntype == 0
        p = addr + 24
        for i in range(used):
            child = struct.unpack_from("<Q", raw, p + 8)[0] + base
            if level > 0:
                walk(child)
            else:
                assert raw[child:child + 4] == b"SNOD"
                nsym = struct.unpack_from("<H", raw, child + 6)[0]
                prev = None
                for j in range(nsym):
                    noff, ohdr, _ct, _ = struct.unpack_from("<QQII", raw, child + 8 + 40 * j)
                    nm = name_at(noff)
                    assert prev is None or prev.encode() < nm.encode(), "symbol table entries must be sorted"
                    prev = nm
                    out[nm] = dataset(ohdr + base)
            p += 16

    def dataset(addr):
        msgs = dict(messages(addr))
        sp = msgs[0x0001]
        ver, rank, flags = struct.unpack_from("<BBB", sp, 0)
        assert ver == 1 and (flags & ~1) == 0                # bit 0: maximum dimensions follow the current ones
        dims = struct.unpack_from("<%dQ" % rank, sp, 8)
        dt = _read_datatype(msgs[0x0003])
        lay = msgs[0x0008]
        count = int(np.prod(dims)) if rank else 1
        if lay[0] == 3:
            lv, lclass, daddr, dsize = struct.unpack_from("<BBQQ", lay, 0)
        else:                                                # libhdf5 1.6: version, dimensionality, class, 5 reserved, address, 32-bit sizes
            assert lay[0] in (1, 2)
            ndim, lclass = lay[1], lay[2]
            daddr = struct.unpack_from("<Q", lay, 8)[0]
            dsize = int(np.prod(struct.unpack_from("<%dI" % ndim, lay, 16)))
        assert lclass == 1
        assert dsize == count * dt.itemsize
        if dsize == 0:
            return np.zeros(dims, dtype=dt)
        daddr += base
        assert daddr % 8 == 0
        return np.frombuffer(raw, dtype=dt, count=count, offset=daddr).reshape(dims).copy()

    walk(btree_addr)
    return out
