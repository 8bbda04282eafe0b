"""ctypes wrapper of the C restatement (oracle/dlrm_oracle.c).  TEST INFRASTRUCTURE ONLY --
see the header of dlrm_oracle.c.  Builds the shared object with `make` on first use."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libdlrm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "dlrm_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libdlrm_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _table_ptrs(tables: Sequence[np.ndarray]):
    for t in tables:
        assert t.dtype == np.float32 and t.flags.c_contiguous
    return (C.c_void_p * len(tables))(*[t.ctypes.data for t in tables])


def _idx64(idx) -> np.ndarray:
    """[ntab][B][P] int64, 0-based."""
    if isinstance(idx, np.ndarray):
        a = idx
    else:
        a = np.stack([np.asarray(i).reshape(np.asarray(i).shape[0], -1) for i in idx])
    if a.ndim == 2:
        a = a[:, :, None]
    return np.ascontiguousarray(a, dtype=np.int64)


def lookup(tables: Sequence[np.ndarray], idx, slot0: int = 0, nthreads: int = 0, out: np.ndarray = None) -> np.ndarray:
    ix = _idx64(idx)
    ntab, B, P = ix.shape
    D = tables[0].shape[1]
    slots = slot0 + ntab
    if out is None:
        out = np.zeros((B, slots, D), dtype=np.float32)
    lib().oracle_lookup(_table_ptrs(tables), ntab, D, _p(ix), B, P, _p(out), slots, slot0,
                        nthreads or max_threads())
    return out


def interaction_fwd(T: np.ndarray, pad_to_mul: int = 1, nthreads: int = 0, out: np.ndarray = None) -> np.ndarray:
    T = np.ascontiguousarray(T, dtype=np.float32)
    B, F, d = T.shape
    width = (d + F * (F - 1) // 2 + pad_to_mul - 1) // pad_to_mul * pad_to_mul
    if out is None:
        out = np.empty((B, width), dtype=np.float32)
    lib().oracle_interaction_fwd(_p(T), B, F, d, pad_to_mul, _p(out), nthreads or max_threads())
    return out


def interaction_bwd(dOut: np.ndarray, T: np.ndarray, pad_to_mul: int = 1, nthreads: int = 0, dT=None, dx=None):
    T = np.ascontiguousarray(T, dtype=np.float32)
    dOut = np.ascontiguousarray(dOut, dtype=np.float32)
    B, F, d = T.shape
    if dT is None:
        dT = np.empty_like(T)
    if dx is None:
        dx = np.empty((B, d), dtype=np.float32)
    lib().oracle_interaction_bwd(_p(dOut), _p(T), B, F, d, pad_to_mul, _p(dT), _p(dx), nthreads or max_threads())
    return dx, dT


def sparse_sgd(tables: List[np.ndarray], idx, dT: np.ndarray, slot0: int, lr: float, nthreads: int = 0) -> None:
    ix = _idx64(idx)
    ntab, B, P = ix.shape
    D = tables[0].shape[1]
    dT = np.ascontiguousarray(dT, dtype=np.float32)
    slots = dT.shape[1]
    lib().oracle_sparse_sgd(_table_ptrs(tables), ntab, D, _p(ix), B, P, _p(dT), slots, slot0,
                            C.c_float(lr), nthreads or max_threads())


def init_uniform(rows: int, D: int, seed: int, nthreads: int = 0) -> np.ndarray:
    """[rows][D] table of U(-1/sqrt(rows), 1/sqrt(rows)) values, filled in parallel (CPU-arm setup)."""
    t = np.empty((rows, D), dtype=np.float32)
    lib().oracle_init_uniform(_p(t), C.c_int64(rows), D, C.c_uint64(seed), nthreads or max_threads())
    return t
