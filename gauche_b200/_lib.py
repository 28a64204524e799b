"""ctypes binding of the C ABI declared in ``include/gauche_b200.h``.

The shared library ``csrc/libgauche_b200.so`` is built in-tree for sm_100a by
``gauche_b200/csrc/Makefile`` (``__graft_entry__.build()`` drives it).  There is no
fallback: if the library is missing, or no CUDA device is present, every compute entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC_DIR, "libgauche_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gauche_b200.h")

# enum gk_dtype
GK_F32, GK_F64, GK_I64, GK_I32, GK_U8, GK_F16, GK_BF16, GK_I8, GK_I16, GK_BOOL = range(10)
GK_FLAG_NONBINARY = 1
GK_FLAG_NONU8 = 2
GK_Q_ABI_VERSION, GK_Q_SM_COUNT, GK_Q_CC, GK_Q_UMMA_TILE_M, GK_Q_UMMA_TILE_N, GK_Q_K_ALIGN = range(6)

_i64, _i32, _p, _dbl, _f32 = C.c_int64, C.c_int, C.c_void_p, C.c_double, C.c_float

# name -> (restype, argtypes); must list every symbol of include/gauche_b200.h
SIGNATURES = {
    "gk_last_error": (C.c_char_p, []),
    "gk_query": (_i32, [_i32, C.POINTER(_i64)]),
    "gk_pack_classify": (_i32, [_p, _i32, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _p, _p]),
    "gk_tanimoto_gram_popc": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_tanimoto_gram_u8": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_tanimoto_gram_bits": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_tanimoto_gram_hybrid": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_tanimoto_rowdot_hybrid": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _p, _dbl, _f32, _p, _i64, _p, _p]),
    "gk_permute_bits_kblock": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "gk_tanimoto_gram_umma2": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _i32, _dbl, _i32, _p]),
    "gk_host_pack_bits": (_i32, [_p, _i32, _i64, _i64, _i64, _p, _i64, _p, C.POINTER(_i32), _i32]),
    "gk_bits_popcount": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "gk_expand_bits_u8": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "gk_expand_bits_e2m1": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "gk_tanimoto_gram_mxf4": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_reduce_partials": (_i32, [_p, _i64, _i64, _i64, _f32, _p, _p]),
    "gk_expand_thermometer_e2m1": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _i64, _p]),
    "gk_block_occupancy_words": (_i64, [_i64, _i64, _i32]),
    "gk_block_occupancy": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _p]),
    "gk_minmax_gram_mxf4_sparse": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _p, _p, _i64, _i32, _dbl, _i32, _p]),
    "gk_minmax_gram_mxf4_sym": (_i32, [_p, _i64, _p, _i64, _i64, _p, _p, _p, _i64, _i32, _dbl, _i32, _p]),
    "gk_minmax_gram_mxf4": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_minmax_gram_u8": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_tanimoto_gram_real": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_minmax_gram_real": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_tanimoto_rowdot_slabs": (_i64, [_i64, _i32]),
    "gk_tanimoto_rowdot": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i32, _p, _dbl, _f32, _p, _i64, _p, _p]),
    "gk_tanimoto_bwd_coeff": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _i32, _dbl, _p]),
    "gk_minmax_bwd": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _i64, _i32, _dbl, _p]),
    "gk_sibling_gram": (_i32, [_i32, _i32, _p, _i64, _p, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i32, _p, _i64, _i32, _dbl, _p]),
    "gk_rowsum_real": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _p]),
    "gk_sibling_gram_real": (_i32, [_i32, _p, _i64, _p, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i32, _p, _i64, _i32, _dbl, _p]),
    "gk_rowdot_f32": (_i32, [_p, _i64, _i64, _i64, _p, _f32, _p, _p]),
    "gk_topk_workspace": (_i64, [_i64, _i32]),
    "gk_topk_f32": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _i64, _p]),
    "gk_selftest_fdiv_f32": (_i32, [_p, _p, _p, _p, _i64, _p]),
    "gk_selftest_fdiv_f64": (_i32, [_p, _p, _p, _p, _i64, _p]),
}

_lock = threading.Lock()
_lib = None


class GaucheB200Error(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status."""


def build(verbose: bool = False) -> str:
    """Compile ``libgauche_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j", "8"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout[-4000:])
        print(proc.stderr[-4000:])
    if proc.returncode != 0:
        raise RuntimeError("building libgauche_b200.so failed (see output above)")
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library and attach signatures.  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GaucheB200Error(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C gauche_b200/csrc`). "
                "gauche_b200 has no CPU or PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    msg = load().gk_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        kind = "argument error" if rc < 0 else "CUDA error"
        raise GaucheB200Error(f"{what} failed with {kind} {rc}: {last_error()}")


def query(what: int) -> int:
    out = _i64(0)
    check(load().gk_query(what, C.byref(out)), "gk_query")
    return int(out.value)
