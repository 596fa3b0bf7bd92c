"""Thin tensor-level wrappers over the C ABI (pointers + current CUDA stream; torch is only
the allocator and the stream provider).  Every function launches exactly one kernel of
libni_b200.so unless stated; nothing here computes on the host."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import NI_BF16, NI_F16, NI_F32, NI_F64, NI_MAX_GEN, NI_MAX_TERMS, NiError, NiStepDesc, check

DTYPE_CODE = {torch.float32: NI_F32, torch.float16: NI_F16, torch.bfloat16: NI_BF16, torch.float64: NI_F64}


def _code(dt: torch.dtype) -> int:
    try:
        return DTYPE_CODE[dt]
    except KeyError:
        raise NiError(f"dtype {dt} is not supported by libni_b200") from None


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise NiError(f"{what} must be a CUDA tensor (got {t.device}); there is no CPU fallback")
    if not t.is_contiguous():
        raise NiError(f"{what} must be contiguous")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def weighted_sum_tensors(coeffs: Sequence[float], tensors: Sequence[torch.Tensor], *, out_dtype: Optional[torch.dtype] = None,
                         scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = scale * sum_t coeffs[t] * tensors[t]  (ni_weighted_sum).  Rows longer than NI_MAX_TERMS
    are not needed by any drop-in (a python list of >512 tensors), and raise."""
    n = len(tensors)
    if n == 0:
        raise NiError("weighted_sum_tensors needs at least one tensor")
    if len(coeffs) < n:
        raise NiError(f"{n} tensors but only {len(coeffs)} coefficients")
    t0 = tensors[0]
    for i, t in enumerate(tensors):
        _require_cuda(t, f"tensors[{i}]")
        if t.dtype != t0.dtype or t.shape != t0.shape or t.device != t0.device:
            raise NiError("all tensors of a weighted sum must share dtype, shape and device")
    if out is None:
        out = torch.empty_like(t0, dtype=out_dtype or t0.dtype)
    else:
        _require_cuda(out, "out")
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    cs = (C.c_double * n)(*[float(c) for c in coeffs[:n]])
    with torch.cuda.device(t0.device):
        check(_lib.lib().ni_weighted_sum(ptrs, cs, n, out.data_ptr(), t0.numel(), _code(t0.dtype), _code(out.dtype),
                                         float(scale), stream_ptr(t0.device)), "ni_weighted_sum")
    return out


def philox_normal(shape, *, seed: int, tensor_id: int, elem_offset: int = 0, dtype=torch.float32, device="cuda",
                  out: Optional[torch.Tensor] = None, elem_offset_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """N(0,1) tensor of the noise contract (include/ni_b200.h); identical values for any sharding
    as long as each shard passes its own global `elem_offset`.  elem_offset_dev: a 1-element int64 CUDA tensor whose
    value is added to elem_offset when the kernel runs (ni_philox_normal_at)."""
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=device)
    _require_cuda(out, "out")
    with torch.cuda.device(out.device):
        if elem_offset_dev is None:
            check(_lib.lib().ni_philox_normal(out.data_ptr(), out.numel(), _code(out.dtype), seed & (2**64 - 1), tensor_id,
                                              elem_offset, stream_ptr(out.device)), "ni_philox_normal")
        else:
            if not elem_offset_dev.is_cuda or elem_offset_dev.dtype != torch.int64 or elem_offset_dev.numel() != 1:
                raise NiError("elem_offset_dev must be a 1-element int64 CUDA tensor")
            check(_lib.lib().ni_philox_normal_at(out.data_ptr(), out.numel(), _code(out.dtype), seed & (2**64 - 1), tensor_id,
                                                 elem_offset, elem_offset_dev.data_ptr(), stream_ptr(out.device)), "ni_philox_normal_at")
    return out


def to_pixel_u8(x: torch.Tensor, *, scale: float = 0.5, shift: float = 0.5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NCHW float -> NHWC uint8 with the reference's truncating cast (src/CIFAR10NaturalInference.py:212-216)."""
    _require_cuda(x, "x")
    if x.dim() != 4:
        raise NiError("to_pixel_u8 expects NCHW")
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().ni_to_pixel_u8(x.data_ptr(), _code(x.dtype), out.data_ptr(), n, c, h, w, scale, shift,
                                        stream_ptr(x.device)), "ni_to_pixel_u8")
    return out


class StepLaunch:
    """One prepared ``ni_step`` launch: owns the ctypes descriptor and its host tables."""

    __slots__ = ("desc", "_ptrs", "_coeffs")

    def __init__(self, *, numel, per_sample, dtype, out_dtype=None, has_x0=True, x_in=0, out0=0, out1=0,
                 out_sample_stride=None, a=0.0, b0=0.0, b1=0.0, x0_dst=0, c_x0=0.0, c_xin=0.0,
                 terms=(), gens=(), seed=0, elem_offset=0, accumulate=False, x_next=0, x_next_lp=0, lp_dtype=NI_BF16, sumsq=0,
                 bias=0.0, pixels_u8=0, px_scale=0.5, px_shift=0.5, px_channels=0, elem_offset_dev=0):
        """terms: iterable of (device_ptr, coeff); gens: iterable of (tensor_id, coeff, dst_ptr_or_0)."""
        terms = list(terms)
        gens = list(gens)
        if len(terms) > NI_MAX_TERMS:
            raise NiError(f"{len(terms)} stored terms > NI_MAX_TERMS; chain launches with accumulate=True")
        if len(gens) > NI_MAX_GEN:
            raise NiError(f"{len(gens)} generated terms > NI_MAX_GEN")
        d = NiStepDesc()
        d.numel, d.per_sample = int(numel), int(per_sample)
        d.dtype = dtype
        d.out_dtype = dtype if out_dtype is None else out_dtype
        d.has_x0 = 1 if has_x0 else 0
        d.x_in, d.out0, d.out1 = x_in or None, out0 or None, out1 or None
        d.out_sample_stride = int(per_sample if out_sample_stride is None else out_sample_stride)
        d.a, d.b0, d.b1 = float(a), float(b0), float(b1)
        d.x0_dst = x0_dst or None
        d.c_x0 = float(c_x0)
        d.c_xin = float(c_xin)
        n = len(terms)
        self._ptrs = (C.c_void_p * max(n, 1))(*[p for p, _ in terms])
        self._coeffs = (C.c_float * max(n, 1))(*[float(c) for _, c in terms])
        d.n_terms = n
        d.term_ptrs_host = C.cast(self._ptrs, C.POINTER(C.c_void_p))
        d.term_coeffs_host = C.cast(self._coeffs, C.POINTER(C.c_float))
        d.n_gen = len(gens)
        for i, (tid, c, dst) in enumerate(gens):
            d.gen_tensor_ids[i] = int(tid)
            d.gen_coeffs[i] = float(c)
            d.gen_dst[i] = dst or None
        d.philox_seed = int(seed) & (2**64 - 1)
        d.elem_offset = int(elem_offset)
        d.elem_offset_dev = elem_offset_dev or None
        d.accumulate = 1 if accumulate else 0
        d.x_next = x_next or None
        d.x_next_lp = x_next_lp or None
        d.lp_dtype = lp_dtype
        d.sumsq = sumsq or None
        d.bias = float(bias)
        d.pixels_u8 = pixels_u8 or None
        d.px_scale, d.px_shift, d.px_channels = float(px_scale), float(px_shift), int(px_channels)
        self.desc = d

    def flavour(self) -> int:
        """1 = streaming loads, 0 = L2-friendly loads: what ni_step picks for this launch on the current device
        (ni_step_flavour; launches nothing)."""
        rc = _lib.lib().ni_step_flavour(C.byref(self.desc))
        if rc < 0:
            check(rc, "ni_step_flavour")
        return rc

    def launch(self, stream: int):
        rc = _lib.lib().ni_step(C.byref(self.desc), stream)
        if rc != 0:
            check(rc, "ni_step")


def fused_step(*, x_in: Optional[torch.Tensor], outs: Sequence[torch.Tensor], a: float, b: Sequence[float],
               c_x0: float, terms: Sequence[tuple], gens: Sequence[tuple] = (), c_xin: float = 0.0, seed: int = 0, elem_offset: int = 0,
               per_sample: Optional[int] = None, out_sample_stride: Optional[int] = None, keep_x0: bool = True,
               keep_gen: Sequence[bool] = (), lp_dtype: Optional[torch.dtype] = None, want_sumsq: bool = False,
               state_dtype: Optional[torch.dtype] = None, shape=None, device=None):
    """Functional one-shot form of ni_step for tests and ad-hoc use (allocates its outputs).

    terms: [(coeff, tensor)], gens: [(coeff, tensor_id)].  Returns dict with x_next and, when
    requested, x0 / gen tensors / x_next_lp / sumsq."""
    ref = x_in if x_in is not None else (terms[0][1] if terms else None)
    if ref is None:
        if shape is None or device is None or state_dtype is None:
            raise NiError("fused_step without tensors needs shape, device and state_dtype")
        dev, dt, shp = torch.device(device), state_dtype, tuple(shape)
    else:
        dev, dt, shp = ref.device, ref.dtype, tuple(ref.shape if shape is None else shape)
    numel = int(np.prod(shp))
    if per_sample is None:
        per_sample = numel // shp[0] if len(shp) > 1 else numel
    for i, (_, t) in enumerate(terms):
        _require_cuda(t, f"terms[{i}]")
        if t.dtype != dt:
            raise NiError("terms must share the state dtype")
    has_x0 = len(outs) > 0
    res = {}
    x_next = torch.empty(shp, dtype=dt, device=dev)
    x0 = torch.empty(shp, dtype=dt, device=dev) if (has_x0 and keep_x0) else None
    gen_dst = []
    for i, _ in enumerate(gens):
        keep = i < len(keep_gen) and keep_gen[i]
        gen_dst.append(torch.empty(shp, dtype=dt, device=dev) if keep else None)
    lp = torch.empty(shp, dtype=lp_dtype, device=dev) if lp_dtype is not None else None
    sumsq = torch.zeros(numel // per_sample, dtype=torch.float32, device=dev) if want_sumsq else None
    for i, o in enumerate(outs):
        _require_cuda(o, f"outs[{i}]")
    bb = list(b) + [0.0, 0.0]
    L = StepLaunch(numel=numel, per_sample=per_sample, dtype=_code(dt), out_dtype=_code(outs[0].dtype) if has_x0 else None,
                   has_x0=has_x0, x_in=x_in.data_ptr() if x_in is not None else 0,
                   out0=outs[0].data_ptr() if has_x0 else 0, out1=outs[1].data_ptr() if len(outs) > 1 else 0,
                   out_sample_stride=out_sample_stride, a=a, b0=bb[0], b1=bb[1], x0_dst=x0.data_ptr() if x0 is not None else 0,
                   c_x0=c_x0, c_xin=c_xin, terms=[(t.data_ptr(), c) for c, t in terms],
                   gens=[(tid, c, gen_dst[i].data_ptr() if gen_dst[i] is not None else 0) for i, (c, tid) in enumerate(gens)],
                   seed=seed, elem_offset=elem_offset, x_next=x_next.data_ptr(), x_next_lp=lp.data_ptr() if lp is not None else 0,
                   lp_dtype=_code(lp_dtype) if lp_dtype is not None else NI_BF16, sumsq=sumsq.data_ptr() if sumsq is not None else 0)
    with torch.cuda.device(dev):
        L.launch(stream_ptr(dev))
    res.update(x_next=x_next, x0=x0, gen=gen_dst, x_next_lp=lp, sumsq=sumsq)
    return res
