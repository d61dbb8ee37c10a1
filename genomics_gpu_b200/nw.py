"""ctypes binding of the full-matrix linear-gap Needleman-Wunsch kernel (include/nw_b200.h; replaces the suite's
Non-CDP/NW nw_gpu0..3).  No fallback: the functions live in libgasal_b200.so and need a CUDA device."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import capi

DEFAULT_SCORES = (1, -1, -1, -1)     # MATCH, MISMATCH, INSERTION, DELETION (NW/common.h:5-8)
NW_SYMBOLS = ["nw_gpu_b200", "nwx_matrix_async", "nwx_matrix_host", "nwx_checksum", "nwx_kernel_launches", "nwx_release", "nwx_set_variant"]


class Scores(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("insertion", C.c_int32), ("deletion", C.c_int32)]


_bound = False


def _lib() -> C.CDLL:
    global _bound
    lib = capi.load()
    if not _bound:
        lib.nw_gpu_b200.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
        lib.nwx_matrix_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(Scores), C.c_void_p]; lib.nwx_matrix_async.restype = C.c_int
        lib.nwx_matrix_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(Scores), C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
        lib.nwx_matrix_host.restype = C.c_int
        lib.nwx_checksum.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]; lib.nwx_checksum.restype = C.c_uint64
        lib.nwx_kernel_launches.restype = C.c_uint64
        lib.nwx_set_variant.argtypes = [C.c_int]
        _bound = True
    return lib


def nw_matrix(reference: np.ndarray, query: np.ndarray, scores: Tuple[int, int, int, int] = DEFAULT_SCORES, *, want_matrix: bool = True,
              want_checksum: bool = False):
    """Host-buffer call (H2D, kernel, optional D2H of the N x N int32 matrix).  Returns (matrix or None, ms[h2d, kernel, d2h], checksum or None)."""
    reference = np.ascontiguousarray(reference, dtype=np.uint8); query = np.ascontiguousarray(query, dtype=np.uint8)
    n = int(reference.size)
    if query.size != n:
        raise ValueError("reference and query must have the same length (NW/main.cu uses one N)")
    out = np.empty((n, n), dtype=np.int32) if want_matrix else None
    ms = (C.c_float * 3)()
    h = C.c_uint64(0)
    sc = Scores(*scores)
    rc = _lib().nwx_matrix_host(reference.ctypes.data, query.ctypes.data, out.ctypes.data if out is not None else None, n, C.byref(sc), ms,
                                C.byref(h) if want_checksum else None)
    if rc != 0:
        raise RuntimeError("nwx_matrix_host failed")
    return out, (ms[0], ms[1], ms[2]), (int(h.value) if want_checksum else None)


def set_variant(v: int) -> None:
    _lib().nwx_set_variant(v)


def release() -> None:
    _lib().nwx_release()
