"""Minimal HDF5 writer / reader for the trajectory container of SURVEY.md 8(f) N4.

The reference writes its trajectories with OpenMM's / MDTraj's `HDF5Reporter` (`Simulation/Simulator.py:364-401`) and
reads them back with `mdtraj.load` (`Analysis/analysis.py:2032-2037`).  Neither h5py, PyTables nor libhdf5 exists in this
image, so this module lays the bytes out itself, following the HDF5 File Format Specification 1.1 ("earliest" library
version -- what libhdf5 itself writes by default): version-0 superblock, a root group stored as symbol table (version-1
B-tree + local heap + one symbol-table node), version-1 object headers, contiguous datasets of little-endian IEEE floats,
two's-complement integers or fixed-length strings, and version-1 attribute messages (scalar or 1-D).  No chunking,
compression, nested groups or variable-length types.

UNPINNED: without libhdf5 here the only check is the round trip through `read()` below (an independent parser of the same
structures) plus structural assertions on signatures, sizes and addresses (tests/test_trajectory.py); DESIGN.md 3 lists it
with the other third-party formats that cannot be checked in this environment.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INTERNAL_K = 16, 16          # symbol-table node holds 2 * _LEAF_K entries
_HEAP_FREE_NULL = 1                    # libhdf5's on-disk "end of free list" marker (H5HL_FREE_NULL)


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------ messages
def _datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits = (8, 23) if dt.itemsize == 4 else (11, 52)
        head = struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize)      # version 1, class 1; LE, implied msb
        prop = struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return head + prop
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)   # class 0; LE, signed bit
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)                          # class 3; null-terminated ASCII
    raise TypeError(f"unsupported dtype {dt}")


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _attribute(name: str, value) -> bytes:
    if isinstance(value, (str, bytes)):
        raw = value.encode("ascii") if isinstance(value, str) else value
        arr = np.array(raw + b"\0", dtype=f"S{len(raw) + 1}")
    else:
        arr = np.ascontiguousarray(value)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
    nm = name.encode("ascii") + b"\0"
    dt, ds = _datatype(arr.dtype), _dataspace(arr.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()
    return _message(0x000C, body)


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


# ------------------------------------------------------------------------------------------------ writer
def write(path: str, datasets: Dict[str, np.ndarray], attrs: Dict[str, object] | None = None,
          dataset_attrs: Dict[str, Dict[str, object]] | None = None) -> None:
    """One flat file: root-group attributes `attrs`, datasets `datasets` (name -> array) with `dataset_attrs[name]`."""
    attrs, dataset_attrs = dict(attrs or {}), dict(dataset_attrs or {})
    names = sorted(datasets, key=lambda s: s.encode("ascii"))          # symbol-table entries are ordered by name
    if len(names) > 2 * _LEAF_K:
        raise ValueError(f"at most {2 * _LEAF_K} datasets")
    arrays = {}
    for n in names:
        a = np.ascontiguousarray(datasets[n])
        if a.dtype.kind == "U":
            a = a.astype("S")
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        arrays[n] = a
    # local heap data segment: "" at offset 0, then the names, then one free block
    heap, name_off = bytearray(8), {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode("ascii") + b"\0")
    heap_used = len(heap)
    heap += struct.pack("<QQ", _HEAP_FREE_NULL, 32) + bytes(16)        # free block: next = end of list, size = 32
    # ---- addresses
    pos = 96                                                            # after the superblock
    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", 0, 0))] + [_attribute(k, v) for k, v in attrs.items()])
    root_addr, pos = pos, pos + len(root_hdr)
    btree_addr, pos = pos, pos + 24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8
    heap_addr, pos = pos, pos + 32
    heap_data_addr, pos = pos, pos + len(heap)
    snod_addr, pos = pos, pos + 8 + 2 * _LEAF_K * 40
    hdr_addr, hdr_len = {}, {}
    for n in names:
        a = arrays[n]
        msgs = [_message(1, _dataspace(a.shape)), _message(3, _datatype(a.dtype)), _message(5, struct.pack("<BBBB", 2, 1, 0, 0)),
                _message(8, struct.pack("<BBQQ", 3, 1, 0, a.nbytes))] + [_attribute(k, v) for k, v in dataset_attrs.get(n, {}).items()]
        hdr_addr[n], hdr_len[n] = pos, len(_object_header(msgs))
        pos += hdr_len[n]
    data_addr = {}
    for n in names:
        pos += -pos % 8
        data_addr[n], pos = pos, pos + arrays[n].nbytes
    eof = pos
    # ---- bytes
    out = bytearray()
    out += SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
    out += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    out += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)   # root symbol-table entry
    assert len(out) == 96
    out += _object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + [_attribute(k, v) for k, v in attrs.items()])
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
    bt += struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]] if names else 0)
    out += bt + bytes(24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8 - len(bt))
    out += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), heap_used, heap_data_addr)
    out += heap
    sn = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for n in names:
        sn += struct.pack("<QQII16x", name_off[n], hdr_addr[n], 0, 0)
    out += sn + bytes(8 + 2 * _LEAF_K * 40 - len(sn))
    for n in names:
        a = arrays[n]
        msgs = [_message(1, _dataspace(a.shape)), _message(3, _datatype(a.dtype)), _message(5, struct.pack("<BBBB", 2, 1, 0, 0)),
                _message(8, struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes))] + [_attribute(k, v) for k, v in dataset_attrs.get(n, {}).items()]
        assert len(out) == hdr_addr[n]
        out += _object_header(msgs)
    with open(path, "wb") as f:
        f.write(out)
        for n in names:
            f.write(bytes(data_addr[n] - f.tell()))
            a = arrays[n]
            if a.ndim and a.nbytes:      # no copy: memory-mapped coordinates are streamed from the page cache
                f.write(memoryview(a.reshape(-1).view(np.uint8)))
            else:
                f.write(a.tobytes())
        assert f.tell() == eof


# ------------------------------------------------------------------------------------------------ reader
def _parse_datatype(b: bytes) -> np.dtype:
    cls, bits0, _, _, size = struct.unpack_from("<BBBBI", b, 0)
    version, cls = cls >> 4, cls & 0x0F
    if version != 1:
        raise ValueError("datatype version")
    order = ">" if bits0 & 1 else "<"
    if cls == 1:
        return np.dtype(f"{order}f{size}")
    if cls == 0:
        return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
    if cls == 3:
        return np.dtype(f"S{size}")
    raise ValueError(f"datatype class {cls}")


def _parse_dataspace(b: bytes) -> Tuple[int, ...]:
    version, rank, flags = struct.unpack_from("<BBB", b, 0)
    if version != 1:
        raise ValueError("dataspace version")
    return tuple(struct.unpack_from(f"<{rank}Q", b, 8)) if rank else ()


def _read_header(buf: memoryview, addr: int):
    version, nmsg, _, size = struct.unpack_from("<BxHII", buf, addr)
    if version != 1:
        raise ValueError("object header version")
    msgs, p, end = [], addr + 16, addr + 16 + size
    while p < end and len(msgs) < nmsg:
        mtype, msize, _ = struct.unpack_from("<HHB", buf, p)
        msgs.append((mtype, bytes(buf[p + 8:p + 8 + msize])))
        p += 8 + msize
    return msgs


def _parse_attribute(b: bytes):
    version, nlen, tlen, slen = struct.unpack_from("<BxHHH", b, 0)
    if version != 1:
        raise ValueError("attribute version")
    p = 8
    name = b[p:p + nlen].split(b"\0")[0].decode("ascii")
    p += nlen + (-nlen % 8)
    dt = _parse_datatype(b[p:p + tlen])
    p += tlen + (-tlen % 8)
    shape = _parse_dataspace(b[p:p + slen])
    p += slen + (-slen % 8)
    n = int(np.prod(shape)) if shape else 1
    val = np.frombuffer(b, dtype=dt, count=n, offset=p).reshape(shape)
    if dt.kind == "S" and not shape:
        return name, val.item().decode("ascii")
    return name, val.copy()


def read(path: str):
    """-> (datasets {name: array}, root attrs {name: value}, dataset attrs {name: {attr: value}}) of a file this module (or
    libhdf5 with its default 'earliest' format, contiguous datasets in the root group) wrote."""
    with open(path, "rb") as f:
        raw = f.read()
    buf = memoryview(raw)
    if raw[:8] != SIGNATURE:
        raise ValueError("not an HDF5 file")
    sb_version, _, _, _, _, size_off, size_len = struct.unpack_from("<BBBBBBB", buf, 8)
    if sb_version != 0 or size_off != 8 or size_len != 8:
        raise ValueError("unsupported superblock")
    leaf_k, _ = struct.unpack_from("<HH", buf, 16)
    eof = struct.unpack_from("<Q", buf, 40)[0]
    if eof != len(raw):
        raise ValueError("end-of-file address does not match the file size")
    _, root_addr, cache, _, btree_addr, heap_addr = struct.unpack_from("<QQIIQQ", buf, 56)
    root_msgs = _read_header(buf, root_addr)
    attrs = dict(_parse_attribute(d) for t, d in root_msgs if t == 0x000C)
    st = [d for t, d in root_msgs if t == 0x0011]
    if not st or struct.unpack("<QQ", st[0][:16]) != (btree_addr, heap_addr):
        raise ValueError("root group symbol-table message")
    if raw[heap_addr:heap_addr + 4] != b"HEAP":
        raise ValueError("local heap signature")
    heap_size, _, heap_data = struct.unpack_from("<QQQ", buf, heap_addr + 8)

    def name_at(off):
        end = raw.index(b"\0", heap_data + off)
        return raw[heap_data + off:end].decode("ascii")

    entries = []

    def walk(node):
        if raw[node:node + 4] != b"TREE":
            raise ValueError("B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", buf, node + 4)
        for i in range(used):
            child = struct.unpack_from("<Q", buf, node + 24 + 8 + 16 * i)[0]
            if level > 0:
                walk(child)
                continue
            if raw[child:child + 4] != b"SNOD":
                raise ValueError("symbol-table node signature")
            nsym = struct.unpack_from("<H", buf, child + 6)[0]
            for s in range(nsym):
                noff, haddr = struct.unpack_from("<QQ", buf, child + 8 + 40 * s)
                entries.append((name_at(noff), haddr))

    walk(btree_addr)
    datasets, dattrs = {}, {}
    for name, haddr in entries:
        msgs = _read_header(buf, haddr)
        shape = _parse_dataspace([d for t, d in msgs if t == 1][0])
        dt = _parse_datatype([d for t, d in msgs if t == 3][0])
        lay = [d for t, d in msgs if t == 8][0]
        lv, lclass, addr, nbytes = struct.unpack_from("<BBQQ", lay, 0)
        if lv != 3 or lclass != 1:
            raise ValueError("only contiguous version-3 layouts are supported")
        n = int(np.prod(shape)) if shape else 1
        if n * dt.itemsize != nbytes or addr + nbytes > len(raw):
            raise ValueError(f"dataset {name}: size mismatch")
        datasets[name] = np.frombuffer(raw, dtype=dt, count=n, offset=addr).reshape(shape).copy()
        dattrs[name] = dict(_parse_attribute(d) for t, d in msgs if t == 0x000C)
    return datasets, attrs, dattrs
