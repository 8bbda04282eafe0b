"""Minimal pure-Python reader for the PyTorch-reference HDF5 files DLRM.jl validates against.

Replaces the HDF5.jl dependency of the reference's model import
(`src/data/criteo.jl:464-560`, `load_hdf5` / `load_inputs`) for the only container
shape those files use: superblock v0, one root symbol-table group, v1 object headers,
contiguous unfiltered little-endian f32 / i64 datasets.  No libhdf5, no h5py (neither
is in the image).  Anything outside that subset raises ``ValueError`` rather than
guessing.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


def _u(buf: bytes, off: int, n: int) -> int:
    return int.from_bytes(buf[off:off + n], "little")


class _File:
    def __init__(self, buf: bytes):
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise ValueError(f"superblock version {buf[8]} unsupported (only v0)")
        if buf[13] != 8 or buf[14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        self.buf = buf
        # root symbol-table entry starts at byte 56 (after four u64 addresses at 24)
        self.root_header = _u(buf, 56 + 8, 8)

    # -- object header v1 --------------------------------------------------------------
    def messages(self, addr: int):
        buf = self.buf
        if buf[addr] != 1:
            raise ValueError(f"object header version {buf[addr]} unsupported (only v1)")
        nmsg = _u(buf, addr + 2, 2)
        hsize = _u(buf, addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, length = blocks.pop(0)
            end = pos + length
            while pos + 8 <= end and len(out) < nmsg:
                mtype = _u(buf, pos, 2)
                msize = _u(buf, pos + 2, 2)
                body = pos + 8
                if mtype == 0x10:  # continuation
                    blocks.append((_u(buf, body, 8), _u(buf, body + 8, 8)))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # -- group walk --------------------------------------------------------------------
    def root_entries(self) -> Dict[str, int]:
        btree = heap = None
        for mtype, body, _ in self.messages(self.root_header):
            if mtype == 0x11:
                btree = _u(self.buf, body, 8)
                heap = _u(self.buf, body + 8, 8)
        if btree is None:
            raise ValueError("root object has no symbol-table message")
        buf = self.buf
        if buf[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        heap_data = _u(buf, heap + 24, 8)
        entries: Dict[str, int] = {}

        def name_at(off: int) -> str:
            start = heap_data + off
            stop = buf.index(b"\x00", start)
            return buf[start:stop].decode("ascii")

        def walk(node: int):
            if buf[node:node + 4] == b"TREE":
                level = buf[node + 5]
                used = _u(buf, node + 6, 2)
                pos = node + 24  # after sig, type, level, used, two sibling addresses
                for i in range(used):
                    child = _u(buf, pos + 8 + i * 16, 8)  # key, child, key, child, ...
                    walk(child)
                del level
            elif buf[node:node + 4] == b"SNOD":
                count = _u(buf, node + 6, 2)
                pos = node + 8
                for i in range(count):
                    e = pos + 40 * i
                    entries[name_at(_u(buf, e, 8))] = _u(buf, e + 8, 8)
            else:
                raise ValueError("unexpected node signature in group b-tree")

        walk(btree)
        return entries

    # -- dataset -----------------------------------------------------------------------
    def dataset(self, addr: int) -> np.ndarray:
        buf = self.buf
        dims = None
        dtype = None
        data_addr = data_size = None
        for mtype, body, msize in self.messages(addr):
            if mtype == 0x01:
                if buf[body] != 1:
                    raise ValueError("dataspace version unsupported")
                rank = buf[body + 1]
                dims = tuple(_u(buf, body + 8 + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                cls = buf[body] & 0x0F
                size = _u(buf, body + 4, 4)
                if cls == 1 and size == 4:
                    dtype = np.dtype("<f4")
                elif cls == 1 and size == 8:
                    dtype = np.dtype("<f8")
                elif cls == 0 and size == 8:
                    dtype = np.dtype("<i8")
                elif cls == 0 and size == 4:
                    dtype = np.dtype("<i4")
                else:
                    raise ValueError(f"datatype class {cls} size {size} unsupported")
            elif mtype == 0x08:
                if buf[body] != 3 or buf[body + 1] != 1:
                    raise ValueError("only contiguous layout v3 supported")
                data_addr = _u(buf, body + 2, 8)
                data_size = _u(buf, body + 10, 8)
            elif mtype == 0x0B:
                raise ValueError("filtered datasets unsupported")
        if dims is None or dtype is None or data_addr is None:
            raise ValueError("incomplete dataset header")
        n = int(np.prod(dims)) if dims else 1
        if n * dtype.itemsize != data_size:
            raise ValueError("dataset size mismatch")
        arr = np.frombuffer(buf, dtype=dtype, count=n, offset=data_addr)
        return arr.reshape(dims).copy()


def read_hdf5(path: str) -> Dict[str, np.ndarray]:
    """Return every root-level dataset as a numpy array in HDF5/C (= PyTorch) order."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    return {name: f.dataset(addr) for name, addr in sorted(f.root_entries().items())}


__all__ = ["read_hdf5"]
del struct
