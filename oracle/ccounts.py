"""ctypes wrapper of oracle/counts.c (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_counts.so")
_lib = None


def build() -> str:
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pack(x: np.ndarray, kp: int):
    """Dense [rows, D] -> (bits uint32 [rows, kp/32], q8 uint8 [rows, kp], rowsum, rowsq, flags)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    rows, d = x.shape
    bits = np.empty((rows, kp // 32), dtype=np.uint32)
    q8 = np.empty((rows, kp), dtype=np.uint8)
    rowsum = np.empty(rows, dtype=np.int32)
    rowsq = np.empty(rows, dtype=np.int32)
    lib = _load()
    lib.oracle_pack.restype = C.c_int
    flags = lib.oracle_pack(_p(x), C.c_int64(rows), C.c_int64(d), C.c_int64(d), _p(bits),
                            C.c_int64(kp // 32), _p(q8), C.c_int64(kp), _p(rowsum), _p(rowsq))
    return bits, q8, rowsum, rowsq, int(flags)


def _pairwise(fn_name, a, b, k):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.int32)
    fn = getattr(_load(), fn_name)
    fn.restype = None
    fn(_p(a), C.c_int64(a.shape[1]), C.c_int64(a.shape[0]), _p(b), C.c_int64(b.shape[1]),
       C.c_int64(b.shape[0]), C.c_int64(k), _p(out), C.c_int64(b.shape[0]))
    return out


def intersections(bits_a: np.ndarray, bits_b: np.ndarray) -> np.ndarray:
    return _pairwise("oracle_intersections", bits_a.astype(np.uint32), bits_b.astype(np.uint32), bits_a.shape[1])


def dot_u8(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return _pairwise("oracle_dot_u8", a.astype(np.uint8), b.astype(np.uint8), a.shape[1])


def l1_u8(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return _pairwise("oracle_l1_u8", a.astype(np.uint8), b.astype(np.uint8), a.shape[1])
