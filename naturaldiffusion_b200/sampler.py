"""The Natural Inference sampler loop (loop-level drop-in).

What the three reference loops do per step (src/CIFAR10NaturalInference.py:294-304,
src/ValidateNaturalInference.py:349-366, src/SD3NaturalInference.py:201-221) becomes ONE
``ni_step`` launch per step here: model I/O scaling + CFG, append-to-history, the row of A
against the x0 ring, the row of B against stored noise and the freshly drawn noise.

The python lists ``seq_x0`` / ``seq_eps`` / ``seq_xstarts`` of the reference (unbounded, fp64 for
CIFAR) become a ring of fixed device slots sized from the liveness of the matrix columns
(coeffs.build_plan): 4 slots for step_10_weight_42, 5 for step_15_weight_173, 14 for the sharp SD3
table, K for dense first-order rows.

Denoiser protocol: ``denoiser(x, k) -> out`` or ``(out0, out1)``; tensors of shape [B, C', H, W] with
C' >= C (only the first C channels are read -- DiT's learned-sigma half is skipped without a copy).
The denoiser stays torch (north-star); adapters for the reference's three model interfaces are in
``naturaldiffusion_b200.adapters``.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import os

import numpy as np
import torch

from . import _lib
from ._lib import NI_BF16, NI_MAX_TERMS, NiError
from .coeffs import CoeffTriple, StepPlan, build_plan
from .ops import DTYPE_CODE, StepLaunch, philox_normal, stream_ptr


class NaturalInferenceSampler:
    def __init__(self, triple: CoeffTriple, io_scaling: Sequence[Tuple[float, float, float]], batch: int,
                 sample_shape: Sequence[int], *, device="cuda", dtype: torch.dtype = torch.float32, seed: int = 0,
                 eps0: str = "stored", lp_dtype: Optional[torch.dtype] = None, track_sumsq: bool = False,
                 sample_offset: int = 0, keep_all_x0: bool = False, markov="auto", final_scale: float = 1.0, final_bias: float = 0.0,
                 advance: Optional[int] = None):
        """
        triple       coefficient matrices (A, B, node)
        io_scaling   K tuples (a_k, b0_k, b1_k): x0_k = a_k x_k + b0_k out0 + b1_k out1 (coeffs.io_*)
        batch        samples on THIS rank; sample_offset = global index of its first sample, so the
                     Philox noise of a sharded run equals the single-GPU run
        advance      samples the noise index moves forward after every trajectory that drew noise in-kernel (default: `batch`;
                     a run sharded over G ranks passes G*batch so shards never overlap; 0 = repeat the same draws).  Like
                     the reference's torch.randn / randn_like (src/ValidateNaturalInference.py:345,359) every call gets new
                     noise -- including the fresh per-step noise of stochastic matrices when the caller supplies the initial
                     tensor.  The index lives in a DEVICE counter read by the kernels, so a captured CUDA graph also draws
                     new noise on each replay.
        eps0         "stored": the initial noise lives in a slot and is re-read by every row that uses it
                     "regen" : rows regenerate it in-kernel from (seed, tensor 0) -- no slot, no reads
        lp_dtype     also emit x_{k+1} in fp16/bf16 for a reduced-precision denoiser (fp32 state only)
        final_scale, final_bias   output stage of latent models folded into the LAST step: x_K <- x_K*final_scale + final_bias
                     (z/0.18215 of src/ValidateNaturalInference.py:368; z/scaling_factor + shift_factor of
                     src/SD3NaturalInference.py:238).  The scale multiplies the last row's coefficients on the host, free.
        markov       "auto" | True | False.  First-order matrices (DDPM, DDIM, Euler, flow Euler; detected by
                     coeffs.markov_ratios to 1e-12) satisfy row_k = c_k*row_{k-1} + new terms, so the history sum
                     equals c_k*x_k and a step reads nothing but the model output and x_k: O(1) instead of O(k)
                     transfers per step, no ring at all.  Same trajectory as the dense rows up to fp32 rounding
                     (it is the original sampler's own arithmetic).
        """
        if eps0 not in ("stored", "regen"):
            raise NiError("eps0 must be 'stored' or 'regen'")
        if dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise NiError(f"state dtype {dtype} not supported")
        self.triple = triple
        self.K = triple.K
        if len(io_scaling) != self.K:
            raise NiError(f"io_scaling has {len(io_scaling)} entries, matrix has K={self.K}")
        self.io = [tuple(float(v) for v in t) for t in io_scaling]
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NiError("NaturalInferenceSampler runs on CUDA only; there is no CPU fallback")
        _lib.lib()
        self.dtype = dtype
        self.batch = int(batch)
        self.sample_shape = tuple(int(s) for s in sample_shape)
        self.per_sample = int(np.prod(self.sample_shape))
        self.numel = self.batch * self.per_sample
        self.seed = int(seed)
        self.eps0_mode = eps0
        self.lp_dtype = lp_dtype
        self.advance = self.batch if advance is None else int(advance)
        # the Philox element offset of this shard: a device counter (all descriptors carry its address, their host offset is 0)
        self._counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._graphs = {}
        self._relay_peer: Optional[torch.device] = None
        self.set_sample_offset(sample_offset)
        self.final_scale, self.final_bias = float(final_scale), float(final_bias)
        from .coeffs import markov_ratios
        use_markov = (markov_ratios(triple) is not None) if markov == "auto" else bool(markov)
        self.plan: StepPlan = build_plan(triple, keep_all_x0=keep_all_x0, markov=use_markov)
        p = self.plan
        # one slab: X ping-pong (2) + eps0 (1) + x0 ring + eps ring
        n_buf = 3 + p.n_x0_slots + p.n_eps_slots
        self._slab = torch.empty((n_buf, self.numel), dtype=dtype, device=self.device)
        self._X = [self._slab[0], self._slab[1]]
        self._eps0 = self._slab[2]
        self._x0_slots = [self._slab[3 + i] for i in range(p.n_x0_slots)]
        self._eps_slots = [self._slab[3 + p.n_x0_slots + i] for i in range(p.n_eps_slots)]
        self._lp = [torch.empty(self.numel, dtype=lp_dtype, device=self.device) for _ in range(2)] if lp_dtype else None
        self.sumsq = torch.zeros((self.K, self.batch), dtype=torch.float32, device=self.device) if track_sumsq else None
        self._launches: Optional[List[List[StepLaunch]]] = None
        self._launch_key = None
        self._launch_cache = {}
        # NVTX ranges around sample() and each fused step (timeline tools: nsys / ncu --nvtx); off unless NI_NVTX=1
        self.nvtx = os.environ.get("NI_NVTX", "") == "1"
        self.kernel_launches_per_trajectory = sum(p.launches(k, eps0 == "stored") for k in range(self.K))

    def set_sample_offset(self, sample_offset: int):
        """Re-target this sampler at batch `[sample_offset, sample_offset + B)` of the global sample index space.  In-kernel
        noise is keyed by the global element index, so a run split into batches (or ranks) of any size draws the same
        samples.  Stream-ordered (a device fill); prepared descriptors and captured graphs stay valid."""
        self.elem_offset = int(sample_offset) * self.per_sample  # host mirror of the device counter
        with torch.cuda.device(self.device):
            self._counter.fill_(self.elem_offset)

    def _advance_noise(self, st: int):
        """After a trajectory that drew noise in-kernel: move the device counter (inside the stream / graph) and its host mirror."""
        if self.advance:
            delta = self.advance * self.per_sample
            _lib.check(_lib.lib().ni_counter_add(self._counter.data_ptr(), delta, st), "ni_counter_add")
            self.elem_offset += delta

    # ------------------------------------------------------------------ views
    def full_shape(self):
        return (self.batch,) + self.sample_shape

    def x0_slot(self, j: int) -> Optional[torch.Tensor]:
        s = self.plan.x0_slot_of[j]
        return None if s < 0 else self._x0_slots[s].view(self.full_shape())

    def state_bytes(self) -> int:
        return self._slab.numel() * self._slab.element_size()

    # ------------------------------------------------------------------ launch preparation
    def _prepare(self, x_init_ptr: int, eps0_ptr: int, fresh_ptrs: Optional[Sequence[int]], out_ptr: int, stored0: bool, pix_ptr: int = 0):
        key = (x_init_ptr, eps0_ptr, tuple(fresh_ptrs) if fresh_ptrs is not None else None, out_ptr, stored0, pix_ptr)
        if self._launches is not None and key == self._launch_key:
            return
        if key in self._launch_cache:  # e.g. the two alternating staging buffers of sample_host_many
            self._launches, self._launch_key = self._launch_cache[key], key
            return
        p, code = self.plan, DTYPE_CODE[self.dtype]
        launches: List[List[StepLaunch]] = []
        for k, s in enumerate(p.steps):
            x_in = x_init_ptr if k == 0 else self._X[k % 2].data_ptr()
            x_next = self._X[(k + 1) % 2].data_ptr()
            if k == self.K - 1 and out_ptr:
                x_next = out_ptr
            terms = [(self._x0_slots[p.x0_slot_of[j]].data_ptr(), c) for j, c in s.hist]
            gens = []
            for j, c in s.eps:
                if j == 0:
                    if stored0:
                        terms.append((eps0_ptr, c))
                    else:
                        gens.append((0, c, 0))
                elif fresh_ptrs is not None:
                    terms.append((fresh_ptrs[j - 1], c))
                else:
                    terms.append((self._eps_slots[p.eps_slot_of[j]].data_ptr(), c))
            if s.fresh is not None:
                if fresh_ptrs is not None:
                    if s.fresh != 0.0:
                        terms.append((fresh_ptrs[k], s.fresh))
                else:
                    dst = self._eps_slots[s.fresh_slot].data_ptr() if s.keep_fresh else 0
                    gens.append((k + 1, s.fresh, dst))
            a, b0, b1 = self.io[k]
            c_x0, c_xin, bias = s.c_x0, s.c_xin, 0.0
            pix = 0
            if k == self.K - 1:  # output stage folded into the last row
                fs = self.final_scale
                if fs != 1.0:
                    terms = [(p_, c * fs) for p_, c in terms]
                    gens = [(tid, c * fs, dst) for tid, c, dst in gens]
                    c_x0, c_xin = c_x0 * fs, c_xin * fs
                bias = self.final_bias
                pix = pix_ptr
            common = dict(numel=self.numel, per_sample=self.per_sample, dtype=code, seed=self.seed, elem_offset=0,
                          elem_offset_dev=self._counter.data_ptr())
            chunks = [terms[i:i + NI_MAX_TERMS] for i in range(0, max(len(terms), 1), NI_MAX_TERMS)]
            if pix and len(chunks) == 1:
                x_next = 0  # the uint8 image replaces x_K (a chained >512-term row still needs x_K to accumulate into)
            row = []
            for ci, chunk in enumerate(chunks):
                last = ci == len(chunks) - 1
                row.append(StepLaunch(
                    **common, has_x0=ci == 0, x_in=x_in if ci == 0 else 0, a=a, b0=b0, b1=b1,
                    x0_dst=(self._x0_slots[s.x0_slot].data_ptr() if (s.keep_x0 and ci == 0) else 0), c_x0=c_x0,
                    c_xin=c_xin if ci == 0 else 0.0, bias=bias if last else 0.0,
                    pixels_u8=pix if last else 0, px_channels=self.sample_shape[0],
                    terms=chunk, gens=gens if ci == 0 else (), accumulate=ci > 0, x_next=x_next,
                    x_next_lp=(self._lp[(k + 1) % 2].data_ptr() if (self._lp is not None and last) else 0),
                    lp_dtype=DTYPE_CODE[self.lp_dtype] if self.lp_dtype else NI_BF16,
                    sumsq=(self.sumsq[k].data_ptr() if (self.sumsq is not None and last) else 0)))
            launches.append(row)
        if len(self._launch_cache) >= 8:
            self._launch_cache.pop(next(iter(self._launch_cache)))
        self._launch_cache[key] = launches
        self._launches, self._launch_key = launches, key

    def load_flavours(self) -> List[int]:
        """Per step, the load flavour the prepared launches get (1 streaming / 0 L2-friendly; after a first sample())."""
        if self._launches is None:
            raise NiError("load_flavours() needs prepared launches: call sample() once")
        return [row[0].flavour() for row in self._launches]

    def _check_out(self, o: torch.Tensor, k: int):
        if not o.is_cuda:
            raise NiError(f"denoiser output at step {k} must be a CUDA tensor")
        if self.batch == 0:
            return
        if o.dim() < 2 or o.shape[0] != self.batch or o.numel() % self.batch != 0 or o.numel() // self.batch < self.per_sample:
            raise NiError(f"denoiser output at step {k} has shape {tuple(o.shape)}; expected [B={self.batch}, >= {self.per_sample} elements]")

    def step(self, k: int, outs, stream: Optional[int] = None):
        """Launch step k on model output(s) `outs` (after _prepare)."""
        if isinstance(outs, torch.Tensor):
            outs = (outs,)
        # a non-contiguous output (channels_last nets, einsum views) is compacted by torch first: the kernel
        # addresses [B, C', H, W] row-major
        outs = tuple(o if (o is None or o.is_contiguous()) else o.contiguous() for o in outs)
        o0 = outs[0]
        self._check_out(o0, k)
        row = self._launches[k]
        d = row[0].desc
        d.out0 = o0.data_ptr()
        d.out_dtype = DTYPE_CODE[o0.dtype]
        d.out_sample_stride = o0.numel() // self.batch if self.batch else self.per_sample
        if len(outs) > 1 and outs[1] is not None:
            self._check_out(outs[1], k)
            if outs[1].dtype != o0.dtype or outs[1].shape != o0.shape:
                raise NiError("both denoiser outputs must share dtype and shape")
            d.out1 = outs[1].data_ptr()
        else:
            d.out1 = None
            if self.io[k][2] != 0.0:
                raise NiError(f"io_scaling[{k}] has b1 != 0 but the denoiser returned one tensor")
        st = stream_ptr(self.device) if stream is None else stream
        for L in row:
            L.launch(st)

    # ------------------------------------------------------------------ the loop
    @torch.no_grad()
    def sample(self, denoiser: Callable, noise: Optional[torch.Tensor] = None, fresh_noise: Optional[Sequence[torch.Tensor]] = None,
               out: Optional[torch.Tensor] = None, record: bool = False, pixels_out: Optional[torch.Tensor] = None):
        """Run the K-step trajectory.  Returns x_K (a view of internal state unless `out` is given;
        valid until the next call).  With record=True returns (x_K, trace) where trace[k] has clones
        of x0_k and x_{k+1} (needs keep_all_x0=True).  With pixels_out ([B,H,W,C] uint8 CUDA tensor) the last step emits
        the NHWC uint8 image directly (fused output stage, src/CIFAR10NaturalInference.py:308-309) and that tensor is
        returned instead of x_K."""
        if record and any(s < 0 for s in self.plan.x0_slot_of):
            raise NiError("record=True needs a sampler built with keep_all_x0=True")
        shape = self.full_shape()
        with torch.cuda.device(self.device):
            if noise is not None:
                if noise.shape != shape or noise.dtype != self.dtype or not noise.is_cuda or not noise.is_contiguous():
                    raise NiError(f"noise must be a contiguous CUDA {self.dtype} tensor of shape {shape}")
                x_init = eps0 = noise
            else:
                tgt = self._eps0 if self.eps0_mode == "stored" else self._X[0]
                x_init = eps0 = philox_normal(shape, seed=self.seed, tensor_id=0, elem_offset=0, elem_offset_dev=self._counter, out=tgt.view(shape))
            fresh_ptrs = None
            if fresh_noise is not None:
                if len(fresh_noise) != self.K:
                    raise NiError(f"fresh_noise needs K={self.K} tensors")
                for t in fresh_noise:
                    if t.shape != shape or t.dtype != self.dtype or not t.is_cuda or not t.is_contiguous():
                        raise NiError("fresh_noise tensors must match the state shape/dtype and be contiguous CUDA tensors")
                fresh_ptrs = [t.data_ptr() for t in fresh_noise]
            if out is not None and (out.shape != shape or out.dtype != self.dtype or not out.is_cuda or not out.is_contiguous()):
                raise NiError("out must match the state shape/dtype and be a contiguous CUDA tensor")
            # caller-provided noise is always read through its pointer (it need not be the Philox tensor)
            stored0 = self.eps0_mode == "stored" or noise is not None
            if pixels_out is not None:
                if len(self.sample_shape) != 3:
                    raise NiError("pixels_out needs a (C,H,W) sample shape")
                c_, h_, w_ = self.sample_shape
                if (pixels_out.shape != (self.batch, h_, w_, c_) or pixels_out.dtype != torch.uint8 or not pixels_out.is_cuda
                        or not pixels_out.is_contiguous() or record or out is not None or len(self.plan.steps) == 0):
                    raise NiError("pixels_out must be a contiguous CUDA uint8 [B,H,W,C] tensor (and excludes out= / record=)")
            self._prepare(x_init.data_ptr(), eps0.data_ptr(), fresh_ptrs, out.data_ptr() if out is not None else 0, stored0,
                          pixels_out.data_ptr() if pixels_out is not None else 0)
            if self.sumsq is not None:
                self.sumsq.zero_()
            st = stream_ptr(self.device)
            trace = []
            x = x_init.view(shape)
            nvtx = torch.cuda.nvtx if self.nvtx else None
            if nvtx:
                nvtx.range_push(f"ni.sample K={self.K} B={self.batch}")
            for k in range(self.K):
                x_model = x if (self._lp is None or k == 0) else self._lp[k % 2].view(shape)
                if nvtx:
                    nvtx.range_push(f"denoiser k={k}")
                outs = denoiser(x_model, k)
                if nvtx:
                    nvtx.range_pop()
                    nvtx.range_push(f"ni_step k={k}")
                self.step(k, outs, st)
                if nvtx:
                    nvtx.range_pop()
                x = (out if (k == self.K - 1 and out is not None) else self._X[(k + 1) % 2]).view(shape)
                if record:
                    trace.append(dict(x0=self.x0_slot(k).clone(), x_next=x.clone()))
            if noise is None or any(L.desc.n_gen for row in self._launches for L in row):
                self._advance_noise(st)
            if nvtx:
                nvtx.range_pop()
        if pixels_out is not None:
            return pixels_out
        return (x, trace) if record else x

    # ------------------------------------------------------------------ CUDA graph of the whole trajectory
    @torch.no_grad()
    def capture(self, denoiser: Callable, noise: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                pixels_out: Optional[torch.Tensor] = None):
        """Capture the K steps (denoiser included) in one CUDA graph.  `noise` (if given) is a static input buffer the
        caller refills between replays; otherwise noise is generated inside the graph.  The noise index is a device counter
        advanced INSIDE the graph, so every replay draws the next batch's noise (set_sample_offset rewinds it).  Several
        graphs (one per buffer set) can coexist; replay() replays the last one captured, replay(g) a given one."""
        with torch.cuda.device(self.device):
            start = self.elem_offset
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.sample(denoiser, noise=noise, out=out, pixels_out=pixels_out)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            before = self.elem_offset
            with torch.cuda.graph(g):
                res = self.sample(denoiser, noise=noise, out=out, pixels_out=pixels_out)
            g._ni_delta = self.elem_offset - before  # what one replay adds to the counter
            g._ni_out = res
            self._graphs[(noise.data_ptr() if noise is not None else 0, out.data_ptr() if out is not None else 0,
                          pixels_out.data_ptr() if pixels_out is not None else 0)] = g
            self._graph = g
            self.elem_offset = start  # capture executed nothing; the warm-up runs are rewound
            self._counter.fill_(start)
        return g

    def replay(self, graph=None) -> torch.Tensor:
        g = graph if graph is not None else getattr(self, "_graph", None)
        if g is None:
            raise NiError("capture() first")
        g.replay()
        self.elem_offset += g._ni_delta
        return g._ni_out

    # ------------------------------------------------------------------ host-buffer entry (end-to-end)
    def set_host_relay(self, peer: Optional[int]):
        """Route the host copies of `sample_host_many` through GPU `peer`: host <-> peer over the PEER's PCIe link, peer <-> this
        GPU over NVLink (one extra 0.9 TB/s hop).  For boxes where this GPU reaches host memory across the socket
        interconnect and an NVLink peer sits next to it -- e.g. a single-NUMA-node VM on a two-socket HGX board, where four
        ranks share ~76 GB/s to host memory and the other four have 177 GB/s (profiles/r02_host_copy_probe_n8.json);
        `hostutil.choose_host_relay` measures whether it pays and picks the peer.  None = direct copies (default).  Results do
        not depend on the route.  The peer only runs copy-engine work for this process (a second CUDA context on it)."""
        if peer is None:
            new = None
        else:
            peer = int(peer)
            own = self.device.index if self.device.index is not None else torch.cuda.current_device()
            if peer == own or peer < 0 or peer >= torch.cuda.device_count():
                raise NiError(f"host relay peer must be another visible GPU (got {peer}, own {own})")
            if not torch.cuda.can_device_access_peer(own, peer):
                raise NiError(f"GPU {own} has no peer access to GPU {peer}")
            new = torch.device("cuda", peer)
        if new != self._relay_peer and hasattr(self, "_stage"):
            # staging buffers, streams and per-buffer graphs are rebuilt for the new route; copies of an earlier call may
            # still be running on the side streams (on both devices), so drain them before the buffers go back to the allocator
            torch.cuda.synchronize(self.device)
            if self._relay_peer is not None:
                torch.cuda.synchronize(self._relay_peer)
            del self._stage
        self._relay_peer = new

    @torch.no_grad()
    def sample_host(self, denoiser: Callable, noise_host: torch.Tensor, out_host: torch.Tensor, pixels: bool = False):
        """End-to-end call with HOST buffers: pinned fp noise in, result out (fp state, or NHWC uint8
        pixels when pixels=True, fused output stage of src/CIFAR10NaturalInference.py:308-309).
        Copies are asynchronous on the current stream; the caller synchronises."""
        shape = self.full_shape()
        if noise_host.device.type != "cpu" or noise_host.shape != shape or noise_host.dtype != self.dtype:
            raise NiError("noise_host must be a CPU tensor matching the state shape/dtype")
        # always staged in the eps_0 slot (never in the X ping-pong buffers, which step 1 overwrites while later rows
        # still read eps_0 through this pointer)
        dev_noise = self._eps0.view(shape)
        dev_noise.copy_(noise_host, non_blocking=True)
        if pixels:
            if len(self.sample_shape) != 3:
                raise NiError("pixels=True needs a (C,H,W) sample shape")
            if not hasattr(self, "_pix"):
                b, (c, h, w) = self.batch, self.sample_shape
                self._pix = torch.empty((b, h, w, c), dtype=torch.uint8, device=self.device)
            x = self.sample(denoiser, noise=dev_noise, pixels_out=self._pix)
        else:
            x = self.sample(denoiser, noise=dev_noise)
        out_host.copy_(x, non_blocking=True)
        return out_host

    @torch.no_grad()
    def sample_host_many(self, denoiser: Callable, noise_hosts: Optional[Sequence[torch.Tensor]], out_hosts: Sequence[torch.Tensor],
                         pixels: bool = False, first_sample: Optional[int] = None, graph: bool = False):
        """Pipelined end-to-end over many batches with HOST buffers (the reference generates 100 batches of 500,
        src/CIFAR10NaturalInference.py:288-309): the H2D copy of batch i+1 and the D2H copy of batch i-1 run on
        their own streams while batch i computes; two device staging buffers per direction.  Returns after
        enqueueing everything; the caller synchronises (torch.cuda.synchronize or the returned event).
        noise_hosts=None: like the reference, draw the noise on the device (in-kernel Philox keyed by the global sample
        index) -- nothing but the results crosses PCIe.  first_sample (optional) rewinds the noise index first; batch i
        then draws samples [first_sample + i*advance, + B).  graph=True replays one captured CUDA graph per batch (two
        graphs, one per staging-buffer parity) instead of K ctypes launches: the denoiser must be capturable."""
        n = len(out_hosts)
        if noise_hosts is not None and len(noise_hosts) != n:
            raise NiError("need one output buffer per noise batch")
        shape = self.full_shape()
        dev = self.device
        if not hasattr(self, "_stage"):
            if pixels:
                if len(self.sample_shape) != 3:
                    raise NiError("pixels=True needs a (C,H,W) sample shape")
                c, h, w = self.sample_shape
                out_shape, out_dtype = (self.batch, h, w, c), torch.uint8
            else:
                out_shape, out_dtype = shape, self.dtype
            self._stage = dict(
                noise=[torch.empty(shape, dtype=self.dtype, device=dev) for _ in range(2)],
                out=[torch.empty(out_shape, dtype=out_dtype, device=dev) for _ in range(2)],
                pixels=pixels, h2d=torch.cuda.Stream(device=dev), d2h=torch.cuda.Stream(device=dev), graphs={})
            peer = self._relay_peer
            if peer is not None:  # relay route: staging buffers and copy streams on the peer, a push stream here
                self._stage.update(
                    r_noise=[torch.empty(shape, dtype=self.dtype, device=peer) for _ in range(2)] if noise_hosts is not None else None,
                    r_out=[torch.empty(out_shape, dtype=out_dtype, device=peer) for _ in range(2)],
                    r_h2d=torch.cuda.Stream(device=peer), r_d2h=torch.cuda.Stream(device=peer), push=torch.cuda.Stream(device=dev))
        st = self._stage
        if st["pixels"] != pixels:
            raise NiError("sample_host_many was first used with a different `pixels` setting on this sampler")
        if first_sample is not None:
            self.set_sample_offset(first_sample)
        main = torch.cuda.current_stream(dev)
        if graph:
            for par in range(min(2, n)):
                key = (par, noise_hosts is not None, id(denoiser))
                if key not in st["graphs"]:
                    nb = st["noise"][par] if noise_hosts is not None else None
                    kw = dict(pixels_out=st["out"][par]) if pixels else dict(out=st["out"][par])
                    st["graphs"][key] = (self.capture(denoiser, noise=nb, **kw), denoiser)  # keeps the denoiser alive with its graph
        relay = self._relay_peer is not None
        if relay and noise_hosts is not None and st["r_noise"] is None:
            st["r_noise"] = [torch.empty(shape, dtype=self.dtype, device=self._relay_peer) for _ in range(2)]
        c_done, d_done, ob_free = [None] * n, [None] * n, [None] * n
        for i in range(n):
            nb, ob = st["noise"][i % 2], st["out"][i % 2]
            if noise_hosts is not None:
                nh = noise_hosts[i]
                if nh.device.type != "cpu" or nh.shape != shape or nh.dtype != self.dtype:
                    raise NiError("noise_hosts[i] must be CPU tensors matching the state shape/dtype")
                with torch.cuda.stream(st["h2d"]):
                    if i >= 2:
                        st["h2d"].wait_event(c_done[i - 2])      # batch i-2 no longer reads this noise buffer
                    else:
                        st["h2d"].wait_stream(main)
                    if relay:
                        # host -> peer staging over the peer's PCIe link, then peer -> here over NVLink.  torch runs a
                        # cross-device copy on the SOURCE device's current stream (r_h2d) after the destination device's
                        # current stream (h2d, which holds the wait above) and makes the latter wait for it.
                        with torch.cuda.stream(st["r_h2d"]):
                            st["r_noise"][i % 2].copy_(nh, non_blocking=True)
                            nb.copy_(st["r_noise"][i % 2], non_blocking=True)
                    else:
                        nb.copy_(nh, non_blocking=True)
                    h_done = torch.cuda.Event()
                    h_done.record(st["h2d"])
                main.wait_event(h_done)
            else:
                nb = None
            if i >= 2:
                main.wait_event(ob_free[i - 2])               # batch i-2's result has left this output buffer
            if graph:
                self.replay(st["graphs"][(i % 2, noise_hosts is not None, id(denoiser))][0])
            elif pixels:
                self.sample(denoiser, noise=nb, pixels_out=ob)
            else:
                self.sample(denoiser, noise=nb, out=ob)
            c_done[i] = torch.cuda.Event()
            c_done[i].record(main)
            if relay:
                with torch.cuda.stream(st["push"]):
                    st["push"].wait_event(c_done[i])
                    with torch.cuda.stream(st["r_d2h"]):
                        # here -> peer staging over NVLink (runs on `push`, after the peer's earlier D2H out of that buffer),
                        # then peer -> host over the peer's PCIe link on r_d2h
                        st["r_out"][i % 2].copy_(ob, non_blocking=True)
                        out_hosts[i].copy_(st["r_out"][i % 2], non_blocking=True)
                        d_done[i] = torch.cuda.Event()
                        d_done[i].record(st["r_d2h"])
                    ob_free[i] = torch.cuda.Event()
                    ob_free[i].record(st["push"])             # the output buffer is free once it has been pushed
            else:
                with torch.cuda.stream(st["d2h"]):
                    st["d2h"].wait_event(c_done[i])
                    out_hosts[i].copy_(ob, non_blocking=True)
                    d_done[i] = torch.cuda.Event()
                    d_done[i].record(st["d2h"])
                ob_free[i] = d_done[i]
        if n:
            main.wait_event(d_done[-1])  # (also across devices: the relay's D2H runs on the peer)
        return d_done[-1] if n else None


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous batch shard [start, stop) of rank `rank` (samples are independent, so
    there is no collective on the sampling path)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
