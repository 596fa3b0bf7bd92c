"""ctypes binding of libni_b200.so (C ABI: include/ni_b200.h).

There is no CPU or eager-torch fallback: if the shared library is missing, or a
compute entry point is called without a CUDA device, this raises.  Build the
library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``python -m naturaldiffusion_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

NI_F32, NI_F16, NI_BF16, NI_F64 = 0, 1, 2, 3
NI_MAX_TERMS = 512
NI_MAX_GEN = 4
NI_ABI_VERSION = 4

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NI_B200_LIB", os.path.join(_HERE, "libni_b200.so"))


class NiError(RuntimeError):
    pass


class NiStepDesc(C.Structure):
    """Mirror of ``struct NiStepDesc`` (include/ni_b200.h) -- keep field order in sync."""
    _fields_ = [
        ("numel", C.c_int64),
        ("per_sample", C.c_int64),
        ("dtype", C.c_int32),
        ("out_dtype", C.c_int32),
        ("has_x0", C.c_int32),
        ("x_in", C.c_void_p),
        ("out0", C.c_void_p),
        ("out1", C.c_void_p),
        ("out_sample_stride", C.c_int64),
        ("a", C.c_float),
        ("b0", C.c_float),
        ("b1", C.c_float),
        ("x0_dst", C.c_void_p),
        ("c_x0", C.c_float),
        ("c_xin", C.c_float),
        ("n_terms", C.c_int32),
        ("term_ptrs_host", C.POINTER(C.c_void_p)),
        ("term_coeffs_host", C.POINTER(C.c_float)),
        ("n_gen", C.c_int32),
        ("gen_tensor_ids", C.c_uint64 * NI_MAX_GEN),
        ("gen_coeffs", C.c_float * NI_MAX_GEN),
        ("gen_dst", C.c_void_p * NI_MAX_GEN),
        ("philox_seed", C.c_uint64),
        ("elem_offset", C.c_uint64),
        ("elem_offset_dev", C.c_void_p),
        ("accumulate", C.c_int32),
        ("bias", C.c_float),
        ("x_next", C.c_void_p),
        ("x_next_lp", C.c_void_p),
        ("lp_dtype", C.c_int32),
        ("sumsq", C.c_void_p),
        ("pixels_u8", C.c_void_p),
        ("px_scale", C.c_float),
        ("px_shift", C.c_float),
        ("px_channels", C.c_int32),
    ]


_lib = None


def lib():
    """Load (once) and return the CDLL; fail loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NiError(
            f"libni_b200.so not found at {LIB_PATH}: the CUDA extension is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.ni_version.restype = C.c_int
    L.ni_last_error.restype = C.c_char_p
    L.ni_launch_count.restype = C.c_int64
    L.ni_lean_launch_count.restype = C.c_int64
    L.ni_set_option.argtypes = [C.c_char_p, C.c_int]
    L.ni_set_option.restype = C.c_int
    L.ni_step.argtypes = [C.POINTER(NiStepDesc), C.c_void_p]
    L.ni_step.restype = C.c_int
    L.ni_step_flavour.argtypes = [C.POINTER(NiStepDesc)]
    L.ni_step_flavour.restype = C.c_int
    L.ni_weighted_sum.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_double), C.c_int, C.c_void_p, C.c_int64,
                                  C.c_int, C.c_int, C.c_double, C.c_void_p]
    L.ni_weighted_sum.restype = C.c_int
    L.ni_philox_normal.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    L.ni_philox_normal.restype = C.c_int
    L.ni_philox_normal_at.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    L.ni_philox_normal_at.restype = C.c_int
    L.ni_counter_add.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.ni_counter_add.restype = C.c_int
    L.ni_debug_box_muller.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.ni_debug_box_muller.restype = C.c_int
    L.ni_fid_accumulate.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.ni_fid_accumulate.restype = C.c_int
    L.ni_to_pixel_u8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                 C.c_float, C.c_float, C.c_void_p]
    L.ni_to_pixel_u8.restype = C.c_int
    if L.ni_version() != NI_ABI_VERSION:
        raise NiError(f"libni_b200.so ABI {L.ni_version()} != binding ABI {NI_ABI_VERSION}; rebuild")
    _lib = L
    return L


EXPORTED_SYMBOLS = ("ni_version", "ni_last_error", "ni_launch_count", "ni_lean_launch_count", "ni_set_option", "ni_step", "ni_step_flavour",
                    "ni_weighted_sum", "ni_philox_normal", "ni_philox_normal_at", "ni_counter_add", "ni_debug_box_muller",
                    "ni_to_pixel_u8", "ni_fid_accumulate")


def check(rc: int, what: str = "libni_b200"):
    if rc != 0:
        msg = lib().ni_last_error().decode("utf-8", "replace")
        raise NiError(f"{what} failed (rc={rc}): {msg}")


def set_option(name: str, value: int):
    check(lib().ni_set_option(name.encode(), int(value)), "ni_set_option")


def launch_count() -> int:
    return int(lib().ni_launch_count())


def lean_launch_count() -> int:
    return int(lib().ni_lean_launch_count())
