"""ctypes front-end of oracle/_build/libni_oracle.so (philox_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libni_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "philox_oracle.c")):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        L = C.CDLL(_SO)
        L.ni_oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.ni_oracle_philox_normal_f32.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64]
        L.ni_oracle_weighted_sum_f32.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.c_int, C.c_void_p, C.c_int64]
        L.ni_oracle_box_muller.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def box_muller(ra: np.ndarray, rb: np.ndarray):
    """(za, zb) of the noise contract for Philox words (ra, rb), evaluated in fp64 and rounded to fp32"""
    ra, rb = np.ascontiguousarray(ra, dtype=np.uint32), np.ascontiguousarray(rb, dtype=np.uint32)
    za, zb = np.empty(ra.shape, dtype=np.float32), np.empty(ra.shape, dtype=np.float32)
    lib().ni_oracle_box_muller(ra.ctypes.data, rb.ctypes.data, za.ctypes.data, zb.ctypes.data, ra.size)
    return za, zb


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().ni_oracle_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def normal(shape, seed: int, tensor_id: int, elem_offset: int = 0) -> np.ndarray:
    out = np.empty(shape, dtype=np.float32)
    lib().ni_oracle_philox_normal_f32(out.ctypes.data, out.size, seed & (2**64 - 1), tensor_id, elem_offset)
    return out


def weighted_sum(coeffs, arrays) -> np.ndarray:
    arrays = [np.ascontiguousarray(a, dtype=np.float32) for a in arrays]
    n = len(arrays)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    cs = (C.c_double * n)(*[float(c) for c in coeffs[:n]])
    out = np.empty_like(arrays[0])
    lib().ni_oracle_weighted_sum_f32(ptrs, cs, n, out.ctypes.data, out.size)
    return out
