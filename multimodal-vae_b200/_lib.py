"""ctypes binding of libmvae_b200.so (the C ABI declared in include/mvae_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing it is built with nvcc, and
if that is impossible every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class MvaeError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("dtype", c_int), ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_void_p), ("lda", c_int64), ("a_major", c_int),
        ("B", c_void_p), ("ldb", c_int64), ("b_major", c_int),
        ("C", c_void_p), ("ldc", c_int64), ("c_dtype", c_int),
        ("bias", c_void_p), ("accumulate", c_int),
        ("col_sum", c_void_p), ("col_sumsq", c_void_p), ("rows_per_group", c_int),
        ("block_n", c_int), ("split_k", c_int), ("stages", c_int),
    ]


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if needed) the shared library.  Raises if it cannot be had."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _build.needs_build():
            _build.build()
        lib = C.CDLL(_build.LIB_PATH)
        lib.mvae_last_error.restype = C.c_char_p
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mvae_last_error().decode("utf-8", "replace")
        raise MvaeError("%s failed (rc=%d): %s" % (what or "mvae call", rc, msg))


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
