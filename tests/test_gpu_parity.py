"""GPU parity: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and against the
golden vectors produced by the reference's own functions.  Tolerance for fp32 state: per step
max|d| <= 1e-5 * ||ref||_2 (north-star); most checks here are far tighter and say so."""
import os

import numpy as np
import pytest
import torch

import naturaldiffusion_b200 as ni
from naturaldiffusion_b200 import dropin
from naturaldiffusion_b200.coeffs import CoeffTriple, ddim_x0_coeffs, io_eps_cfg, io_score_vp, io_velocity_cfg
from naturaldiffusion_b200.ops import fused_step, philox_normal, to_pixel_u8, weighted_sum_tensors
from naturaldiffusion_b200.sampler import NaturalInferenceSampler as _Sampler
from oracle import ni_oracle as O
from oracle import philox
from toy_models import ToyEps

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"


def NaturalInferenceSampler(*a, **kw):
    """The tests in this file re-run one sampler and compare runs, so the noise index stays put (advance=0).  The default
    (every call draws new noise, like torch.randn in the reference) is covered by tests/test_gpu_noise_index.py."""
    kw.setdefault("advance", 0)
    return _Sampler(*a, **kw)


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def rel_err(got, ref):
    """north-star metric: max-abs error relative to the tensor norm"""
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    return ((got - ref).abs().max() / ref.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------ noise
def test_library_loaded_is_in_tree():
    from naturaldiffusion_b200 import _lib
    assert os.path.samefile(os.path.dirname(_lib.LIB_PATH), os.path.dirname(_lib.__file__))
    assert ni.lib().ni_version() == _lib.NI_ABI_VERSION == 4


@pytest.mark.parametrize("numel,off", [(4096, 0), (4100, 0), (1000, 4), (1001, 7), (5, 1), (1 << 20, 1 << 33)])
def test_philox_normal_matches_oracle(numel, off):
    got = philox_normal((numel,), seed=888, tensor_id=3, elem_offset=off, device=DEV).cpu().numpy()
    ref = philox.normal((numel,), seed=888, tensor_id=3, elem_offset=off)
    # integer part bit-exact; fp32 logf + SFU sqrt / sin / cos (abs. error 2^-20.9 times the radius) vs fp64 libm
    assert np.abs(got - ref).max() < 6e-6 and np.abs(got - ref).mean() < 4e-7


def test_philox_normal_sharding_and_dtypes():
    full = philox_normal((8, 3, 32, 32), seed=5, tensor_id=0, device=DEV)
    lo = philox_normal((4, 3, 32, 32), seed=5, tensor_id=0, device=DEV)
    hi = philox_normal((4, 3, 32, 32), seed=5, tensor_id=0, elem_offset=4 * 3072, device=DEV)
    assert torch.equal(full, torch.cat([lo, hi]))
    for dt in (torch.float16, torch.bfloat16):
        h = philox_normal((8, 3, 32, 32), seed=5, tensor_id=0, dtype=dt, device=DEV)
        assert torch.equal(h, full.to(dt))
    # unaligned view -> scalar path, same values
    buf = torch.empty(8 * 3072 + 1, device=DEV)
    v = philox_normal((8 * 3072,), seed=5, tensor_id=0, out=buf[1:])
    assert torch.equal(v, full.flatten())


# ------------------------------------------------------------------ weighted sum
@pytest.mark.parametrize("src,dst", [(torch.float32, torch.float32), (torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16),
                                     (torch.float16, torch.float32), (torch.bfloat16, torch.float32), (torch.float64, torch.float32),
                                     (torch.float64, torch.float64)])
@pytest.mark.parametrize("n_terms,numel", [(1, 64), (3, 4099), (9, 12288), (40, 2048), (300, 520)])
def test_weighted_sum_matches_oracle(src, dst, n_terms, numel):
    g = torch.Generator().manual_seed(n_terms * 1000 + numel)
    xs = [torch.randn(numel, generator=g, dtype=torch.float64).to(src) for _ in range(n_terms)]
    cs = (torch.randn(n_terms, generator=g, dtype=torch.float64) / n_terms ** 0.5).tolist()
    got = weighted_sum_tensors(cs, [x.to(DEV) for x in xs], out_dtype=dst, scale=0.75)
    assert got.dtype == dst
    ref = sum(c * x.double() for c, x in zip(cs, xs)) * 0.75
    tol = {torch.float64: 1e-13, torch.float32: 2e-6, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}[dst]
    scale = ref.abs().max().item() + 1.0
    assert (got.cpu().double() - ref).abs().max().item() <= tol * scale


def test_weighted_sum_unaligned_and_empty():
    buf = torch.randn(3, 1025, device=DEV)
    xs = [buf[i, 1:] for i in range(3)]  # 4-byte aligned only
    got = weighted_sum_tensors([0.5, -1.0, 2.0], [x.contiguous() if not x.is_contiguous() else x for x in xs])
    ref = 0.5 * xs[0] + (-1.0) * xs[1] + 2.0 * xs[2]
    assert torch.allclose(got, ref, atol=1e-6)
    e = weighted_sum_tensors([1.0], [torch.empty(0, device=DEV)])
    assert e.numel() == 0


def test_c_weighted_sum_oracle_agrees():
    """the plain-C fp64 restatement and the kernel agree on the same inputs"""
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(3000, generator=g) for _ in range(6)]
    cs = [0.3, -0.2, 0.0, 1.5, -0.7, 0.11]
    ref = philox.weighted_sum(cs, [x.numpy() for x in xs])
    got = weighted_sum_tensors(cs, [x.to(DEV) for x in xs]).cpu().numpy()
    assert np.abs(got - ref).max() < 3e-6


# ------------------------------------------------------------------ fused step vs fp64 oracle
@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", ["plain", "cfg", "cfg_strided", "no_xin", "ragged"])
def test_fused_step_matches_oracle(dt, case):
    g = torch.Generator().manual_seed(["plain", "cfg", "cfg_strided", "no_xin", "ragged"].index(case))
    B, C, H, Wd = (5, 4, 8, 8) if case != "ragged" else (3, 3, 5, 7)
    shape = (B, C, H, Wd)
    mk = lambda *s: torch.randn(*s, generator=g).to(dt)
    x = mk(*shape)
    Cout = 2 * C if case == "cfg_strided" else C
    outs = [mk(B, Cout, H, Wd)] + ([mk(B, Cout, H, Wd)] if case.startswith("cfg") else [])
    hist = [mk(*shape) for _ in range(5)]
    eps = [mk(*shape) for _ in range(2)]
    a = 0.0 if case == "no_xin" else 1.7
    b = [-0.9, 0.4][: len(outs)]
    A_row = [0.3, -0.1, 0.0, 0.25, -0.4, 1.3]
    B_row = [0.6, 0.2]
    terms = [(c, h.to(DEV)) for c, h in zip(A_row[:5], hist) if c != 0] + [(c, e.to(DEV)) for c, e in zip(B_row, eps)]
    res = fused_step(x_in=None if case == "no_xin" else x.to(DEV), outs=[o.to(DEV) for o in outs], a=a, b=b, c_x0=A_row[5],
                     terms=terms, per_sample=C * H * Wd, out_sample_stride=Cout * H * Wd, want_sumsq=True,
                     state_dtype=dt, shape=shape, device=DEV)
    outs_used = [o[:, :C] for o in outs]
    x0_ref, nxt_ref = O.ni_step_f64(a, b, x, outs_used, A_row, [h.to(dt) for h in hist], B_row, eps)
    x0_ref_r = x0_ref.to(dt)  # the kernel rounds x0 to the storage dtype before it enters the sum
    nxt_ref = nxt_ref - A_row[5] * x0_ref + A_row[5] * x0_ref_r.double()
    tol = {torch.float32: 1e-6, torch.float16: 1.5e-3, torch.bfloat16: 1.2e-2}[dt]
    assert (res["x0"].cpu().double() - x0_ref).abs().max() <= tol * (x0_ref.abs().max() + 1)
    assert (res["x_next"].cpu().double() - nxt_ref).abs().max() <= tol * (nxt_ref.abs().max() + 1)
    ss_ref = (res["x_next"].cpu().double() ** 2).sum(dim=(1, 2, 3))
    assert torch.allclose(res["sumsq"].cpu().double(), ss_ref, rtol=1e-5)


def test_fused_step_generated_noise_kept_and_lp():
    shape = (4, 3, 16, 16)
    g = torch.Generator().manual_seed(1)
    x, o = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    res = fused_step(x_in=x.to(DEV), outs=[o.to(DEV)], a=1.1, b=[-0.3], c_x0=0.9, terms=[], gens=[(0.5, 0), (0.25, 7)],
                     seed=42, elem_offset=4 * 768 * 10, keep_gen=[False, True], lp_dtype=torch.bfloat16)
    e0 = torch.from_numpy(philox.normal(shape, seed=42, tensor_id=0, elem_offset=4 * 768 * 10))
    e7 = torch.from_numpy(philox.normal(shape, seed=42, tensor_id=7, elem_offset=4 * 768 * 10))
    assert (res["gen"][1].cpu() - e7).abs().max() < 6e-6 and res["gen"][0] is None
    ref = 0.9 * (1.1 * x.double() - 0.3 * o.double()) + 0.5 * e0.double() + 0.25 * e7.double()
    assert (res["x_next"].cpu().double() - ref).abs().max() < 6e-6
    assert torch.equal(res["x_next_lp"], res["x_next"].to(torch.bfloat16))


@pytest.mark.parametrize("odt", [torch.float16, torch.bfloat16])
def test_fused_step_reduced_precision_denoiser_io(odt):
    """fp32 sampler state next to an fp16/bf16 denoiser: model outputs are read in their own dtype (8-byte vectors) and
    x_{k+1} is emitted in fp32 and, fused, in the denoiser's dtype"""
    g = torch.Generator().manual_seed(7)
    shape = (6, 4, 16, 16)
    x = torch.randn(shape, generator=g)
    o0, o1 = torch.randn(shape, generator=g).to(odt), torch.randn(shape, generator=g).to(odt)
    h = [torch.randn(shape, generator=g) for _ in range(3)]
    res = fused_step(x_in=x.to(DEV), outs=[o0.to(DEV), o1.to(DEV)], a=1.2, b=[-0.8, 0.3], c_x0=0.7,
                     terms=[(0.2, h[0].to(DEV)), (-0.4, h[1].to(DEV)), (0.9, h[2].to(DEV))], lp_dtype=odt)
    x0 = 1.2 * x.double() - 0.8 * o0.double() + 0.3 * o1.double()
    ref = 0.7 * x0 + 0.2 * h[0].double() - 0.4 * h[1].double() + 0.9 * h[2].double()
    assert (res["x0"].cpu().double() - x0).abs().max() < 2e-6 and (res["x_next"].cpu().double() - ref).abs().max() < 3e-6
    assert torch.equal(res["x_next_lp"], res["x_next"].to(odt))


def test_edge_sizes_empty_batch_and_64bit_indexing(weights_dir):
    """empty shard (a rank with no samples) is a no-op; element indices beyond 2^31 are addressed correctly"""
    triple, s = _c2_sampler(weights_dir, 0)
    assert s.sample(lambda x, k: x).shape == (0, 3, 32, 32)
    assert to_pixel_u8(torch.empty(0, 3, 32, 32, device=DEV)).shape == (0, 32, 32, 3)  # NULL data pointers, still a no-op
    n = (1 << 31) + 4096
    src = torch.empty(n, dtype=torch.float16, device=DEV)
    src[:8] = 1.0
    src[-8:] = 3.0
    src[(1 << 31) - 4:(1 << 31) + 4] = 2.0
    out = weighted_sum_tensors([0.5], [src])
    assert out[:8].eq(0.5).all() and out[-8:].eq(1.5).all() and out[(1 << 31) - 4:(1 << 31) + 4].eq(1.0).all()
    del src, out
    z = philox_normal((8,), seed=1, tensor_id=0, elem_offset=(1 << 40), device=DEV).cpu().numpy()
    assert np.abs(z - philox.normal((8,), seed=1, tensor_id=0, elem_offset=(1 << 40))).max() < 6e-6


def test_fused_step_long_row_chains_with_accumulate():
    """a dense 600-term row (> NI_MAX_TERMS) through the sampler's chunking"""
    from naturaldiffusion_b200.ops import StepLaunch, stream_ptr, DTYPE_CODE
    g = torch.Generator().manual_seed(2)
    n, numel = 600, 4096
    xs = torch.randn(n, numel, generator=g).to(DEV)
    cs = (torch.randn(n, generator=g) / 25).tolist()
    out = torch.empty(numel, device=DEV)
    for ci, lo in enumerate(range(0, n, 512)):
        hi = min(n, lo + 512)
        L = StepLaunch(numel=numel, per_sample=numel, dtype=DTYPE_CODE[torch.float32], has_x0=False,
                       terms=[(xs[i].data_ptr(), cs[i]) for i in range(lo, hi)], accumulate=ci > 0, x_next=out.data_ptr())
        L.launch(stream_ptr(torch.device(DEV)))
    ref = (xs.cpu().double() * torch.tensor(cs, dtype=torch.float64)[:, None]).sum(0)
    assert (out.cpu().double() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape,cout,ncond", [((5, 4, 32, 32), 4, 1), ((3, 4, 32, 32), 8, 2), ((7, 3, 32, 32), 3, 2), ((1, 1, 1, 1000), 1, 1), ((33, 3, 32, 32), 3, 1)])
def test_tma_variant_is_bit_identical_to_direct_loads(dt, shape, cout, ncond):
    """the TMA-staged kernel (cp.async.bulk + mbarrier ring) and the direct-load kernel share the epilogue and the
    accumulation order: same bits, including partial last tiles, strided model outputs, Philox terms and sumsq"""
    from naturaldiffusion_b200 import _lib
    g = torch.Generator().manual_seed(sum(shape) + cout)
    mk = lambda *s: torch.randn(*s, generator=g).to(dt).to(DEV)
    B, C, H, W = shape
    x = mk(*shape)
    outs = [mk(B, cout, H, W) for _ in range(ncond)]
    terms = [(0.1 * (i + 1) * (-1) ** i, mk(*shape)) for i in range(7)]
    kw = dict(x_in=x, outs=outs, a=1.3, b=[-0.7, 0.2][:ncond], c_x0=0.8, terms=terms, gens=[(0.3, 5)], seed=9, keep_gen=[True],
              per_sample=C * H * W, out_sample_stride=cout * H * W, want_sumsq=True)
    try:
        _lib.set_option("variant", 1)
        ref = fused_step(**kw)
        _lib.set_option("variant", 2)
        n0 = ni.launch_count()
        got = fused_step(**kw)
        assert ni.launch_count() == n0 + 1
    finally:
        _lib.set_option("variant", 0)
    for key in ("x_next", "x0"):
        assert torch.equal(got[key], ref[key]), key
    assert torch.equal(got["gen"][0], ref["gen"][0])
    assert torch.allclose(got["sumsq"], ref["sumsq"], rtol=1e-5)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_load_flavours_are_bit_identical(dt, weights_dir):
    """the streaming (plain ld.global) and the L2-friendly (L1::no_allocate) builds of the step kernel differ in cache
    hints only; `auto` picks by the launch footprint vs the L2 size, so small and large launches are both exercised"""
    from naturaldiffusion_b200 import _lib
    g = torch.Generator().manual_seed(11)
    mk = lambda *s: torch.randn(*s, generator=g).to(dt).to(DEV)
    shape = (6, 3, 32, 32)
    kw = dict(x_in=mk(*shape), outs=[mk(*shape), mk(*shape)], a=0.9, b=[-0.4, 0.25], c_x0=0.7, c_xin=0.05,
              terms=[(0.2 * (-1) ** i, mk(*shape)) for i in range(11)], gens=[(0.3, 2)], seed=4, keep_gen=[True], want_sumsq=True)
    res = {}
    try:
        for pol in (1, 2, 0):
            _lib.set_option("load_policy", pol)
            res[pol] = fused_step(**kw)
        with pytest.raises(ni.NiError):
            _lib.set_option("load_policy", 3)
        # a whole C2-sized trajectory (launch footprints 200-400 MB: streaming under `auto`) against the L2-friendly build
        triple = CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz"))
        s = NaturalInferenceSampler(triple, io_score_vp(triple.node), 4096, (3, 32, 32), device=DEV, seed=888)
        den = lambda x, k: torch.tanh(x) * (1 + 0.01 * k)
        _lib.set_option("load_policy", 0)
        auto = s.sample(den).clone()
        _lib.set_option("load_policy", 1)
        assert torch.equal(s.sample(den), auto)
    finally:
        _lib.set_option("load_policy", 0)
    for pol in (2, 0):
        for key in ("x_next", "x0"):
            assert torch.equal(res[pol][key], res[1][key]), (pol, key)
        assert torch.equal(res[pol]["gen"][0], res[1]["gen"][0])


def test_error_paths_are_loud():
    with pytest.raises(ni.NiError):
        weighted_sum_tensors([1.0], [torch.zeros(4)])  # CPU tensor: no fallback
    x = torch.zeros(8, device=DEV)
    with pytest.raises(ni.NiError, match="aliases"):
        weighted_sum_tensors([1.0], [x], out=x)
    with pytest.raises(ni.NiError):
        weighted_sum_tensors([1.0] * 600, [x] * 600)


# ------------------------------------------------------------------ full loops vs reference golden vectors
@pytest.mark.parametrize("name", ["step_5_weight_00", "step_10_weight_42", "step_15_weight_173"])
def test_cifar_loop_matches_reference_golden(golden_dir, weights_dir, name):
    g = _g(golden_dir, "cifar_loop.npz")
    triple = CoeffTriple.from_npz(os.path.join(weights_dir, name + ".npz"))
    noise = torch.from_numpy(g[name + "/noise"]).to(DEV)
    net = ToyEps(3, seed=11)
    ts = triple.node[:, 0]
    s = NaturalInferenceSampler(triple, io_score_vp(triple.node), noise.shape[0], noise.shape[1:], device=DEV, keep_all_x0=True)
    den = lambda x, k: net(x, torch.full((x.shape[0],), float(ts[k]), device=DEV) * 999)
    x, trace = s.sample(den, noise=noise, record=True)
    for k in range(triple.K):
        ref = torch.from_numpy(g[name + "/x_next"][k])
        assert rel_err(trace[k]["x_next"], ref) < 2e-7, f"step {k}"
        assert rel_err(trace[k]["x0"], torch.from_numpy(g[name + "/x0"][k])) < 2e-7
    # ring-buffer version (4/5 live slots instead of all K) gives the same bits
    s2 = NaturalInferenceSampler(triple, io_score_vp(triple.node), noise.shape[0], noise.shape[1:], device=DEV)
    assert s2.plan.n_x0_slots < triple.K
    assert torch.equal(s2.sample(den, noise=noise), x)


@pytest.mark.parametrize("markov", [False, True])
@pytest.mark.parametrize("alg,K", [("ddpm", 24), ("ddim", 24), ("ddpm_sympy", 18), ("ddim", 100)])
def test_validate_loop_matches_reference_golden(golden_dir, alg, K, markov):
    g = _g(golden_dir, "validate_loop.npz")
    m = _g(golden_dir, "reference_matrices.npz")
    fam, key = alg.replace("_sympy", ""), f"{alg}_{K:03d}"
    triple = CoeffTriple(*(m[f"{fam}/{key}/{n}"] for n in ("A", "B", "node")))
    c1, c2, _ = ddim_x0_coeffs(K)
    net = ToyEps(4, seed=23, out_channels=8)
    noise = torch.from_numpy(g[key + "/noise"]).to(DEV)
    fresh = [torch.from_numpy(f).to(DEV) for f in g[key + "/fresh"]]

    def den(z, k):  # full 8-channel outputs; the kernel reads channels [:4] through out_sample_stride
        ts = torch.ones(z.shape[0], dtype=torch.int32, device=DEV) * int(triple.node[k, 0])
        return net(z, ts, 0), net(z, ts, 1)

    s = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, 4.0), noise.shape[0], noise.shape[1:], device=DEV, keep_all_x0=True, markov=markov)
    assert s.plan.markov == markov and (s.plan.n_eps_slots == 0 if markov else True)
    z, trace = s.sample(den, noise=noise, fresh_noise=fresh, record=True)
    # dense rows: the reference NI arithmetic; Markov rows (c_k * x_k stands for the history): the original sampler's
    # arithmetic -- both within the reference's own NI-vs-original gap of the golden vectors
    for k in range(K):
        assert rel_err(trace[k]["x_next"], torch.from_numpy(g[key + "/ni_x_next"][k])) < (5e-6 if markov else 1e-6), f"step {k}"
    assert rel_err(z, torch.from_numpy(g[key + "/original_final"])) < 5e-6  # == the ORIGINAL ddpm/ddim sampler


@pytest.mark.parametrize("wname", ["sd3_step_28_weight", "sd3_step_28_weight_sharp"])
@pytest.mark.parametrize("tag,dt,tol", [("f32", torch.float32, 1e-6), ("f16", torch.float16, 2e-3)])
def test_sd3_loop_matches_reference_golden(golden_dir, weights_dir, wname, tag, dt, tol):
    """fp32 state vs the reference run in fp32; fp16 state (fp32 accumulate) vs the reference's all-fp16
    arithmetic -- the latter differs by fp16 rounding of the reference's running sums (stated tolerance)."""
    g = _g(golden_dir, "sd3_loop.npz")
    sig = g["sigmas"]
    triple = CoeffTriple.from_sd3_csv(os.path.join(weights_dir, wname + ".csv"), sig)
    net = ToyEps(16, seed=5)
    noise = torch.from_numpy(g[f"{wname}/{tag}/noise"]).to(DEV, dt)
    den = lambda x, k: (net(x, 1000 * float(sig[k]), 0), net(x, 1000 * float(sig[k]), 1))
    s = NaturalInferenceSampler(triple, io_velocity_cfg(sig, 7.0), noise.shape[0], noise.shape[1:], device=DEV, dtype=dt, keep_all_x0=True, markov=False)
    out, trace = s.sample(den, noise=noise, record=True)
    if wname == "sd3_step_28_weight":  # the default table is first-order in x0: x_k + eps_0 replace the 27-term history
        sm = NaturalInferenceSampler(triple, io_velocity_cfg(sig, 7.0), noise.shape[0], noise.shape[1:], device=DEV, dtype=dt)
        assert sm.plan.markov and sm.plan.n_x0_slots == 0
        assert rel_err(sm.sample(den, noise=noise), torch.from_numpy(g[f"{wname}/{tag}/out"][27])) < (5e-6 if dt == torch.float32 else 4e-3)
    for k in range(27):
        assert rel_err(trace[k]["x_next"], torch.from_numpy(g[f"{wname}/{tag}/x_in"][k + 1])) < tol, f"step {k}"
    assert rel_err(out, torch.from_numpy(g[f"{wname}/{tag}/out"][27])) < tol


# ------------------------------------------------------------------ drop-ins (reference signatures)
def test_dropin_functions(golden_dir, weights_dir):
    g = _g(golden_dir, "sd3_loop.npz")
    xs = [torch.from_numpy(x).to(DEV) for x in g["fn/xs"]]
    assert rel_err(dropin.weighted_sum(xs, None), torch.from_numpy(g["fn/uniform_mean"])) < 1e-7
    seq = [[float(w), x] for w, x in zip(g["fn/euler_w"], xs)]
    acc, eq = dropin.euler_weighted_sum(seq, 0)
    assert rel_err(acc, torch.from_numpy(g["fn/euler_acc"])) < 1e-7 and rel_err(eq, torch.from_numpy(g["fn/euler_equiv"])) < 1e-7
    assert rel_err(dropin.euler_weighted_sum(seq, 3)[1], torch.from_numpy(g["fn/euler_clip3_equiv"])) < 1e-7
    # device-scalar weights, as the reference passes them (sigma differences on the GPU)
    seq_t = [[torch.tensor(float(w), device=DEV), x] for w, x in zip(g["fn/euler_w"], xs)]
    assert torch.equal(dropin.euler_weighted_sum(seq_t, 0)[1], eq)
    # CIFAR form on an fp64 history, as the unmodified data_fn produces it; zero and -0.0 coefficients
    A, B, node = O.load_triple(os.path.join(weights_dir, "step_10_weight_42.npz"))
    hist64 = [torch.randn(4, 3, 8, 8, dtype=torch.float64) for _ in range(5)]
    ref = O.cifar_weighted_sum(A[4], hist64)
    got = dropin.weighted_sum(A[4], [h.to(DEV) for h in hist64])
    assert got.dtype == torch.float32 and rel_err(got, ref) < 1e-7
    # SD3 form with a table, memoised second call
    W = O.load_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight_sharp.csv"))
    seq7 = [torch.randn(2, 16, 8, 8).to(DEV) for _ in range(7)]
    r1 = dropin.weighted_sum(seq7, W)
    assert rel_err(r1, O.sd3_weighted_sum([t.cpu() for t in seq7], W)) < 1e-6
    assert dropin.weighted_sum(seq7, W) is r1
    seq7[3].add_(1.0)  # an in-place change, or new tensors that happen to reuse the addresses, must not hit the memo
    r2 = dropin.weighted_sum(seq7, W)
    assert r2 is not r1 and rel_err(r2, O.sd3_weighted_sum([t.cpu() for t in seq7], W)) < 1e-6
    fresh = [t.clone() for t in seq7]
    assert dropin.weighted_sum(fresh, W) is not r2


def test_dropin_data_fn_and_install():
    import types
    net = ToyEps(3, seed=11)
    score_fn = O.make_vp_score_fn(lambda x, labels: net(x, labels))
    x = torch.randn(4, 3, 8, 8)
    ref = O.cifar_data_fn(score_fn, x, 0.65, 0.1182, 0.993)
    got = dropin.data_fn(score_fn, x.to(DEV), 0.65, 0.1182, 0.993, DEV)
    assert rel_err(got, ref) < 2e-7
    mod = types.SimpleNamespace(weighted_sum=None, data_fn=None)
    assert set(dropin.install(mod)) == {"weighted_sum", "data_fn"} and mod.weighted_sum is dropin.weighted_sum


# ------------------------------------------------------------------ output stage
def test_fused_output_stage_in_last_step(weights_dir):
    """f3: the last step can emit NHWC uint8 directly (same bytes as ni_to_pixel_u8 on x_K == the reference's
    inverse scaler + to_pixel), and the latent un-scaling x/s + shift folds into the last row"""
    den = lambda x, k: torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x
    triple, s = _c2_sampler(weights_dir, 64)
    x = s.sample(den).clone()
    pix = torch.empty(64, 32, 32, 3, dtype=torch.uint8, device=DEV)
    n0 = ni.launch_count()
    got = s.sample(den, pixels_out=pix)
    assert ni.launch_count() - n0 == triple.K + 1  # K steps + the Philox noise kernel; no separate pixel kernel
    assert torch.equal(got, to_pixel_u8(x)) and np.array_equal(got.cpu().numpy(), O.to_pixel_u8(x.cpu()))
    for variant in (2,):
        from naturaldiffusion_b200 import _lib
        try:
            _lib.set_option("variant", variant)
            assert torch.equal(s.sample(den, pixels_out=torch.empty_like(pix)), got)
        finally:
            _lib.set_option("variant", 0)
    s2 = NaturalInferenceSampler(triple, io_score_vp(triple.node), 64, (3, 32, 32), device=DEV, seed=888, final_scale=1 / 0.18215, final_bias=0.0609)
    assert rel_err(s2.sample(den), x / 0.18215 + 0.0609) < 1e-6


def test_to_pixel_matches_reference_truncation():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(16, 3, 32, 32, generator=g) * 0.8
    x.view(-1)[:6] = torch.tensor([-1.0, 1.0, 0.999999, -3.0, 5.0, 0.0])
    assert np.array_equal(to_pixel_u8(x.to(DEV)).cpu().numpy(), O.to_pixel_u8(x))


# ------------------------------------------------------------------ BASELINE sizes: size-independent properties
def _c2_sampler(weights_dir, batch, **kw):
    triple = CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz"))
    return triple, NaturalInferenceSampler(triple, io_score_vp(triple.node), batch, (3, 32, 32), device=DEV, seed=888, **kw)


def test_full_size_c2_sharding_and_regen_bit_identical(weights_dir):
    """config C2 shape [4096,3,32,32]: (i) two half-batch shards with global Philox offsets == the whole batch,
    (ii) regenerating eps_0 in-kernel == reading the stored tensor, bit for bit."""
    # element-wise denoiser: bit-reproducible for any batch split (a GEMM-based net may pick another algorithm)
    den = lambda x, k: torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x
    _, full = _c2_sampler(weights_dir, 4096)
    ref = full.sample(den).clone()
    _, regen = _c2_sampler(weights_dir, 4096, eps0="regen")
    assert torch.equal(regen.sample(den), ref)
    halves = []
    for r in range(2):
        _, sh = _c2_sampler(weights_dir, 2048, sample_offset=2048 * r)
        halves.append(sh.sample(den).clone())
    assert torch.equal(torch.cat(halves), ref)


def test_host_buffer_paths_match_device_path(weights_dir):
    """end-to-end entry points with HOST buffers (single call and the double-buffered pipeline over many batches)
    return exactly what the device-resident path computes, including the fused uint8 output stage"""
    den = lambda x, k: torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x
    triple, s = _c2_sampler(weights_dir, 256, advance=256)
    g = torch.Generator().manual_seed(4)
    noises = [torch.randn(256, 3, 32, 32, generator=g).pin_memory() for _ in range(5)]
    outs = [torch.empty(256, 32, 32, 3, dtype=torch.uint8).pin_memory() for _ in range(5)]
    s.sample_host_many(den, noises, outs, pixels=True)
    torch.cuda.synchronize()
    refs = [to_pixel_u8(s.sample(den, noise=n.to(DEV))).cpu() for n in noises]
    for o, r in zip(outs, refs):
        assert torch.equal(o, r)
    o1 = torch.empty(256, 32, 32, 3, dtype=torch.uint8).pin_memory()
    s.sample_host(den, noises[3], o1, pixels=True)
    torch.cuda.synchronize()
    assert torch.equal(o1, refs[3])
    # seeded variant: noise drawn on the device, keyed by the global sample index -> equals one big sharded run
    so = [torch.empty(256, 32, 32, 3, dtype=torch.uint8).pin_memory() for _ in range(3)]
    s.sample_host_many(den, None, so, pixels=True, first_sample=512)
    torch.cuda.synchronize()
    _, big = _c2_sampler(weights_dir, 768, sample_offset=512)
    assert torch.equal(torch.cat(so), to_pixel_u8(big.sample(den)).cpu())
    _, s2 = _c2_sampler(weights_dir, 256)
    lat = [torch.empty(256, 3, 32, 32).pin_memory() for _ in range(3)]
    s2.sample_host_many(den, noises[:3], lat, pixels=False)
    torch.cuda.synchronize()
    for n, o in zip(noises, lat):
        assert torch.equal(o, s2.sample(den, noise=n.to(DEV)).cpu())


def test_host_buffer_pipeline_with_a_non_image_sample_shape(weights_dir):
    """the host-buffer pipeline does not assume (C,H,W) samples unless the uint8 pixel stage is asked for: a flat
    1-D sample shape runs through sample_host_many / sample_host and equals the device path; pixels=True refuses it"""
    den = lambda x, k: torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x
    triple = ni.CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz"))
    s = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), 32, (1000,), device=DEV, seed=3)
    g = torch.Generator().manual_seed(9)
    noises = [torch.randn(32, 1000, generator=g).pin_memory() for _ in range(3)]
    outs = [torch.empty(32, 1000).pin_memory() for _ in range(3)]
    s.sample_host_many(den, noises, outs)
    torch.cuda.synchronize()
    for n, o in zip(noises, outs):
        assert torch.equal(o, s.sample(den, noise=n.to(DEV)).cpu())
    o1 = torch.empty(32, 1000).pin_memory()
    s.sample_host(den, noises[1], o1)
    torch.cuda.synchronize()
    assert torch.equal(o1, outs[1])
    s2 = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), 32, (1000,), device=DEV, seed=3)
    with pytest.raises(ni.NiError):
        s2.sample_host_many(den, noises, [torch.empty(32, 1000, dtype=torch.uint8).pin_memory() for _ in range(3)], pixels=True)
    with pytest.raises(ni.NiError):
        s2.sample_host(den, noises[0], torch.empty(32, 1000, dtype=torch.uint8).pin_memory(), pixels=True)


def test_full_size_c2_linearity(weights_dir):
    """with a linear denoiser the whole trajectory is linear in the initial noise"""
    den = lambda x, k: (0.3 + 0.01 * k) * x
    triple, s = _c2_sampler(weights_dir, 4096)
    n1 = philox_normal((4096, 3, 32, 32), seed=1, tensor_id=0, device=DEV)
    n2 = philox_normal((4096, 3, 32, 32), seed=2, tensor_id=0, device=DEV)
    y1 = s.sample(den, noise=n1).clone()
    y2 = s.sample(den, noise=n2).clone()
    y12 = s.sample(den, noise=(n1 + n2))
    assert rel_err(y12, y1 + y2) < 1e-7
    # and equals the closed form  x_K = g * noise  with g from the scalar recursion in fp64
    gk, x0s = 1.0, []
    for k in range(triple.K):
        a, b0, _ = io_score_vp(triple.node)[k]
        x0s.append((a + b0 * (0.3 + 0.01 * k)) * gk)
        gk = sum(triple.A[k, j] * x0s[j] for j in range(k + 1)) + triple.B[k, 0]
    assert rel_err(y1, gk * n1.double()) < 1e-6


def test_full_size_c4_ddpm250_equals_original_sampler():
    """config C4: DDPM ancestral 250 steps, batch 1024 x 4x32x32, CFG 4.0 with two 8-channel model outputs, the
    generated ddpm_250 matrix (dense 250x250 / 250x251 rows) and in-kernel Philox noise.  Checked against
    (i) the reference's NI arithmetic (oracle restatement of src/ValidateNaturalInference.py:343-366: fp32 products,
    fp64 accumulation) and (ii) the ORIGINAL ancestral sampler (:235-250), both run on the same Philox tensors."""
    from toy_models import ToyVPDenoiser
    from naturaldiffusion_b200.generators import ddpm_triple
    K, B = 250, 1024
    triple = ddpm_triple(K)
    c1, c2, _ = ddim_x0_coeffs(K)
    net = ToyVPDenoiser(4, out_channels=8)
    den = lambda z, k: (net(z, triple.node[k, 0], 0), net(z, triple.node[k, 0], 1))
    s = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, 4.0), B, (4, 32, 32), device=DEV, seed=0, markov=False)
    assert s.plan.n_x0_slots >= K - 2 and s.plan.n_eps_slots >= K - 2  # dense rows: everything stays live
    z = s.sample(den).clone()
    del s
    sm = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, 4.0), B, (4, 32, 32), device=DEV, seed=0)  # markov="auto"
    assert sm.plan.markov and sm.plan.n_x0_slots == 0 and sm.plan.n_eps_slots == 0 and sm.plan.total_units(2) == 4 * K
    zm = sm.sample(den).clone()
    noise = philox_normal((B, 4, 32, 32), seed=0, tensor_id=0, device=DEV)
    fresh = [philox_normal((B, 4, 32, 32), seed=0, tensor_id=k + 1, device=DEV) for k in range(K)]
    eps_model = lambda zz, t: tuple(o[:, :4] for o in (net(zz, t, 0), net(zz, t, 1)))
    zn, _ = O.validate_ni_loop(triple.A, triple.B, triple.node, eps_model, noise, fresh)
    zo, _ = O.ddpm_original_loop(K, eps_model, noise, fresh)
    assert float(zo.abs().max()) < 50  # the trajectory is well-conditioned
    assert rel_err(z, zn) < 2e-6
    assert rel_err(zn, zo) < 1e-5 and rel_err(z, zo) < 1e-5
    assert rel_err(zm, zo) < 2e-6 and rel_err(zm, zn) < 1e-5  # the O(1)-reads path IS the original sampler's arithmetic


# ------------------------------------------------------------------ real (random-init) denoisers end to end
def test_c1_ddim10_ncsnpp_ni_equals_original_ddim():
    """config C1: DDIM 10 steps, NCSN++ (61.8 M params, random init, last conv re-initialised), batch 64 x 3x32x32.
    The reference's consistency check (src/ValidateNaturalInference.py:375-391): the ORIGINAL DDIM loop and Natural
    Inference with the matching matrix give the same samples -- here with the NI side on the fused CUDA path."""
    from naturaldiffusion_b200.denoisers import NCSNppVP
    from naturaldiffusion_b200.generators import ddim_triple
    K, B = 10, 64
    torch.manual_seed(0)
    model = NCSNppVP().reinit_output(std=0.02).to(DEV).eval()
    assert abs(sum(p.numel() for p in model.parameters()) - 61.8e6) < 0.1e6
    triple = ddim_triple(K)
    c1, c2, idx = ddim_x0_coeffs(K)
    assert list(idx) == [999, 888, 777, 666, 555, 444, 333, 222, 111, 0] and abs(c1[0] - 157.41046) < 1e-4
    eps_net = lambda z, t: model(z, torch.full((z.shape[0],), float(t), device=DEV))  # discrete label = time index
    den = lambda z, k: eps_net(z, triple.node[k, 0])
    s = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, None), B, (3, 32, 32), device=DEV, seed=0, keep_all_x0=True)
    z, trace = s.sample(den, record=True)
    noise = philox_normal((B, 3, 32, 32), seed=0, tensor_id=0, device=DEV)
    single = lambda zz, t: (eps_net(zz, t), torch.zeros_like(zz))
    zo, otrace = O.ddim_original_loop(K, single, noise, cfg_scale=1.0)
    zn, ntrace = O.validate_ni_loop(triple.A, triple.B, triple.node, single, noise, [torch.zeros_like(noise)] * K, cfg_scale=1.0)
    for k in range(K):
        assert rel_err(trace[k]["x_next"], ntrace[k]["x_next"]) < 1e-5, f"vs reference NI arithmetic, step {k}"
        assert rel_err(trace[k]["x_next"], otrace[k]["x_next"]) < 1e-5, f"vs original DDIM, step {k}"
    assert float((z - noise).abs().mean()) > 0.1  # the network mattered


def test_dit_and_mmdit_adapters_drive_the_sampler():
    """small DiT / MMDiT instances (same code as the XL/2 and SD3-medium shapes) through the CFG adapters:
    two 8-channel outputs read with a sample stride (DiT), fp16 state with two velocity outputs (SD3 loop)."""
    from naturaldiffusion_b200.adapters import dit_cfg_denoiser, mmdit_cfg_denoiser
    from naturaldiffusion_b200.denoisers import DiT, MMDiT
    from naturaldiffusion_b200.generators import ddpm_triple
    torch.manual_seed(0)
    K, B = 18, 4
    dit = DiT(dim=64, depth=2, heads=4).reinit_output(std=0.05).to(DEV).eval()
    triple = ddpm_triple(K)
    c1, c2, _ = ddim_x0_coeffs(K)
    labels = torch.tensor([207, 360, 387, 974], device=DEV)
    den = dit_cfg_denoiser(dit, triple.node, labels)
    s = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, 4.0), B, (4, 32, 32), device=DEV, seed=3)
    z = s.sample(den).clone()
    noise = philox_normal((B, 4, 32, 32), seed=3, tensor_id=0, device=DEV)
    fresh = [philox_normal((B, 4, 32, 32), seed=3, tensor_id=k + 1, device=DEV) for k in range(K)]
    def eps_model(zz, t):
        a, b = dit_cfg_denoiser(dit, triple.node, labels, batched=False)(zz, [int(v) for v in triple.node[:, 0]].index(int(t)))
        return a[:, :4], b[:, :4]
    zo, _ = O.ddpm_original_loop(K, eps_model, noise, fresh)
    assert rel_err(z, zo) < 1e-5

    sig = O.sd3_sigmas()
    W = O.load_sd3_csv(os.path.join(os.path.dirname(os.path.dirname(__file__)), "naturaldiffusion_b200", "data", "weights", "sd3_step_28_weight_sharp.csv"))
    mm = MMDiT(dim=64, depth=2, heads=4, ctx_dim=32, pooled_dim=16, max_grid=32).to(DEV).half().eval()
    ctx, pooled = torch.randn(2, 7, 32, device=DEV).half(), torch.randn(2, 16, device=DEV).half()
    nctx, npooled = torch.randn(2, 7, 32, device=DEV).half(), torch.randn(2, 16, device=DEV).half()
    den3 = mmdit_cfg_denoiser(mm, sig, ctx, pooled, nctx, npooled)
    t3 = CoeffTriple.from_sd3_table(W, sig)
    s3 = NaturalInferenceSampler(t3, io_velocity_cfg(sig, 7.0), 2, (16, 32, 32), device=DEV, dtype=torch.float16, seed=10)
    out = s3.sample(den3).clone()
    n3 = philox_normal((2, 16, 32, 32), seed=10, tensor_id=0, dtype=torch.float16, device=DEV)
    ref, _ = O.sd3_ni_loop(W, sig, lambda x, k: mmdit_cfg_denoiser(mm, sig, ctx, pooled, nctx, npooled, batched=False)(x, k), n3.float())
    assert rel_err(out.float(), ref) < 3e-3  # fp16 state vs the loop evaluated in fp32 on an fp16 network


def test_dpm_solver_pp_2m_matrix_equals_original_multistep_solver():
    """BASELINE config 3's named sampler: DPM-Solver++(2M), 15 steps, quadratic time grid, CIFAR shapes.  Its matrix
    is generated by the coefficient-space tracer (dense 15x15) and run through the fused step; the ORIGINAL multistep
    solver (oracle restatement of deps/dpm_solver_pytorch.py:547-576,796-831) on the same noise gives the same samples."""
    from naturaldiffusion_b200.generators import VPLinearSchedule, dpm_solver_pp_2m_triple
    K, B = 15, 256
    triple = dpm_solver_pp_2m_triple(K)
    ns = VPLinearSchedule()
    ts = triple.node[:, 0]
    io = [(1.0 / ns.alpha(t), -ns.sigma(t) / ns.alpha(t), 0.0) for t in ts[:-1]]  # x0 = (x - sigma*eps)/alpha
    net = ToyEps(3, seed=11, t_scale=0.3)
    eps_model = lambda x, t: float(ns.sigma(t)) * x + 0.1 * net(x, float(t))  # bounded eps-predictor (see ToyVPDenoiser)
    s = NaturalInferenceSampler(triple, io, B, (3, 32, 32), device=DEV, seed=5)
    assert not s.plan.markov and s.plan.n_x0_slots == K - 1
    x = s.sample(lambda z, k: eps_model(z, ts[k]))
    noise = philox_normal((B, 3, 32, 32), seed=5, tensor_id=0, device=DEV)
    xo = O.dpmpp_2m_original_loop(ts, eps_model, noise)
    assert float(xo.abs().max()) < 50
    assert rel_err(x, xo) < 1e-5


@pytest.mark.parametrize("seed", range(12))
def test_random_sparse_matrices_ring_buffer_vs_fp64_oracle(seed):
    """random lower-triangular A / B with random zero patterns (banded, dense, empty columns, stochastic rows), random
    K, ragged shapes, one or two model outputs, stored or regenerated eps_0: the ring-buffer sampler equals the fp64
    common-form recursion step by step"""
    rng = np.random.default_rng(seed)
    K = int(rng.integers(1, 13))
    shape = [(3, 8, 8), (4, 4, 4), (3, 5, 7), (1, 1, 10)][seed % 4]
    B_ = int(rng.integers(1, 6))
    A = np.tril(rng.standard_normal((K, K))) * (rng.random((K, K)) < rng.uniform(0.3, 1.0))
    A[np.arange(K), np.arange(K)] = rng.standard_normal(K) + 2.0
    Bm = np.zeros((K, K + 1))
    stochastic = seed % 3 == 0
    for k in range(K):
        Bm[k, 0] = rng.standard_normal() * (rng.random() < 0.8)
        if stochastic:
            Bm[k, 1:k + 2] = rng.standard_normal(k + 1) * (rng.random(k + 1) < 0.6)
    node = np.stack([np.linspace(1, 0, K + 1), np.linspace(0, 1, K + 1), np.linspace(1, 0, K + 1)], 1)
    triple = CoeffTriple(A, Bm, node)
    m = 1 + seed % 2
    io = [(float(rng.uniform(0.5, 1.5)), float(rng.standard_normal()), float(rng.standard_normal()) if m == 2 else 0.0) for _ in range(K)]
    W1, W2 = rng.standard_normal(2) * 0.3

    def den(x, k):
        o0 = torch.tanh(x * float(W1)) + 0.05 * k
        return o0 if m == 1 else (o0, torch.sin(x * float(W2)))

    eps0 = "regen" if seed % 4 == 1 else "stored"
    s = NaturalInferenceSampler(triple, io, B_, shape, device=DEV, seed=seed, eps0=eps0, keep_all_x0=True, markov=False)
    x, trace = s.sample(den, record=True)
    s_ring = NaturalInferenceSampler(triple, io, B_, shape, device=DEV, seed=seed, eps0=eps0, markov=False)
    assert torch.equal(s_ring.sample(den), x)  # liveness-based ring == keep-everything, bit for bit
    other = NaturalInferenceSampler(triple, io, B_, shape, device=DEV, seed=seed, eps0=("stored" if eps0 == "regen" else "regen"), markov=False)
    assert rel_err(other.sample(den), x) < 1e-6  # stored vs regenerated eps_0: same values, summed in a different position
    full = (B_,) + shape
    eps = [philox_normal(full, seed=seed, tensor_id=j, device=DEV).double() for j in range(K + 1)]
    xk, hist = eps[0], []
    for k in range(K):
        outs = den(xk.float(), k)
        outs = (outs,) if m == 1 else outs
        x0, xk = O.ni_step_f64(io[k][0], io[k][1:1 + m], xk.float(), outs, A[k], hist, Bm[k], eps)
        x0 = x0.float().double()
        hist.append(x0)
        xk = xk.float().double()
        # random rows can cancel three orders of magnitude (|x0| ~ 1e3 -> |x_next| ~ 10): the north-star bound, not tighter
        assert rel_err(trace[k]["x0"], x0) < 2e-6 and rel_err(trace[k]["x_next"], xk) < 1e-5, (seed, k)


def test_deis_tab3_matrix_equals_original_sampler():
    """DEIS tAB3, 15 steps on the quadratic grid (the other sampler BASELINE config 3 names): generated matrix through
    the fused step vs the original exponential-integrator multistep loop (oracle restatement of th_deis)"""
    from naturaldiffusion_b200.generators import VPLinearSchedule, deis_tab_triple
    K, B = 15, 256
    triple = deis_tab_triple(K)
    ns = VPLinearSchedule()
    ts = triple.node[:, 0]
    io = [(1.0 / ns.alpha(t), -ns.sigma(t) / ns.alpha(t), 0.0) for t in ts[:-1]]
    net = ToyEps(3, seed=11, t_scale=0.3)
    eps_model = lambda x, t: float(ns.sigma(t)) * x + 0.1 * net(x, float(t))
    s = NaturalInferenceSampler(triple, io, B, (3, 32, 32), device=DEV, seed=6)
    assert not s.plan.markov
    x = s.sample(lambda z, k: eps_model(z, ts[k]))
    noise = philox_normal((B, 3, 32, 32), seed=6, tensor_id=0, device=DEV)
    xo = O.deis_tab_original_loop(ts, eps_model, noise)
    assert float(xo.abs().max()) < 50 and rel_err(x, xo) < 1e-5


def test_c_abi_from_plain_cpp_without_torch(tmp_path):
    """the boundary is a real C ABI: examples/c_abi_demo.cu (plain C++/CUDA runtime, no Python, no torch) links only
    libni_b200.so, runs a 3-step trajectory with history, Philox noise and the fused uint8 stage, and checks it on the host"""
    import subprocess
    from naturaldiffusion_b200 import build
    exe = build.build_demo()  # prebuilt by __graft_entry__.build(); recompiled only if a source is newer
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libni_b200" in ldd and "torch" not in ldd and "python" not in ldd
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pixel mismatches = 0" in r.stdout


def test_cifar10_pipeline_example_is_batching_invariant():
    """examples/cifar10_pipeline.py (the reference's CIFAR driver on this path: batches -> fused steps -> uint8 images ->
    features -> FID statistics -> all-reduce -> Frechet distance): the statistics do not depend on how the run is cut
    into batches, because noise is keyed by the global sample index"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("cifar10_pipeline", os.path.join(root, "examples", "cifar10_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a = mod.run(samples=96, batch=40, small_model=True, feat_dim=32, quiet=True)  # batches of 40, 40, 16
    b = mod.run(samples=96, batch=96, small_model=True, feat_dim=32, quiet=True)
    assert a["n"] == b["n"] == 96
    assert np.abs(a["mu"] - b["mu"]).max() < 1e-4 and np.abs(a["sigma"] - b["sigma"]).max() < 1e-4
    assert np.isfinite(a["fid"]) and abs(a["fid"] - b["fid"]) < 1e-2 * max(1.0, abs(b["fid"]))


def test_sd3_flow_euler_matrix_equals_vanilla_euler():
    """src/SD3NaturalInference.py:81-154: Natural Inference with the Euler-equivalent weights (sigma_i - sigma_{i+1}) is
    the vanilla flow-matching Euler update (`is_vanilla_update=True`, :126-127).  Here: the exact flow-Euler matrix on the
    SD3 sigma grid through the fused step (first-order path, CFG 7 on two velocity outputs) vs the vanilla Euler loop."""
    from naturaldiffusion_b200.generators import flow_euler_triple
    sig = O.sd3_sigmas().astype(np.float64)
    triple = flow_euler_triple(28, sigmas=sig)
    net = ToyEps(16, seed=5)
    den = lambda x, k: (net(x, 1000 * float(sig[k]), 0), net(x, 1000 * float(sig[k]), 1))
    s = NaturalInferenceSampler(triple, io_velocity_cfg(sig, 7.0), 3, (16, 16, 16), device=DEV, seed=10)
    assert s.plan.markov and s.plan.total_units(2) == 28 * 4
    x = s.sample(den).clone()
    noise = philox_normal((3, 16, 16, 16), seed=10, tensor_id=0, device=DEV)
    ref = O.sd3_euler_original_loop(sig, den, noise)
    assert rel_err(x, ref) < 2e-6
    dense = NaturalInferenceSampler(triple, io_velocity_cfg(sig, 7.0), 3, (16, 16, 16), device=DEV, seed=10, markov=False)
    assert rel_err(dense.sample(den), ref) < 2e-6


def test_sampler_chains_rows_longer_than_512_terms():
    """dense DDPM-300 rows reach 300 + 301 stored terms > NI_MAX_TERMS: the sampler chains two launches per step
    (accumulate=1); result equals the first-order path and the fused uint8 stage still works on a chained last row"""
    from naturaldiffusion_b200.generators import ddpm_triple
    from toy_models import ToyVPDenoiser
    K, B = 300, 4
    triple = ddpm_triple(K)
    c1, c2, _ = ddim_x0_coeffs(K)
    net = ToyVPDenoiser(3)
    den = lambda z, k: net(z, triple.node[k, 0], 0)
    dense = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, None), B, (3, 8, 8), device=DEV, seed=2, markov=False)
    assert max(dense.plan.launches(k) for k in range(K)) == 2 and dense.kernel_launches_per_trajectory > K
    n0 = ni.launch_count()
    xd = dense.sample(den).clone()
    assert ni.launch_count() - n0 == dense.kernel_launches_per_trajectory + 1
    fast = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, None), B, (3, 8, 8), device=DEV, seed=2)
    assert fast.plan.markov and rel_err(xd, fast.sample(den)) < 1e-5
    pix = torch.empty(B, 8, 8, 3, dtype=torch.uint8, device=DEV)
    assert torch.equal(dense.sample(den, pixels_out=pix), to_pixel_u8(xd))


def test_cuda_graph_capture_of_a_whole_trajectory(weights_dir):
    """sampler.capture(): denoiser (torch ops) + K fused steps in one CUDA graph; replays reproduce the eager run and
    follow a refilled static noise buffer"""
    den = lambda x, k: torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x
    triple, s = _c2_sampler(weights_dir, 128)
    noise = philox_normal((128, 3, 32, 32), seed=3, tensor_id=0, device=DEV)
    eager = s.sample(den, noise=noise).clone()
    s.capture(den, noise=noise)
    n0 = ni.launch_count()
    assert torch.equal(s.replay(), eager) and ni.launch_count() == n0  # replays launch from the graph, not through the ABI
    noise.copy_(philox_normal((128, 3, 32, 32), seed=4, tensor_id=0, device=DEV))
    got = s.replay().clone()
    assert torch.equal(got, s.sample(den, noise=noise)) and not torch.equal(got, eager)


def test_presets_build_the_three_reference_configurations(weights_dir):
    """presets.cifar / dit / sd3: the loop-level drop-in in one line per reference script"""
    from naturaldiffusion_b200 import presets
    from naturaldiffusion_b200.denoisers import DiT, MMDiT, NCSNppVP
    from naturaldiffusion_b200.generators import ddim_triple
    torch.manual_seed(0)
    s, wrap = presets.cifar(os.path.join(weights_dir, "step_5_weight_00.npz"), batch=4)
    net = NCSNppVP(nf=32, num_res_blocks=1).reinit_output().to(DEV).eval()
    pix = s.sample(wrap(net), pixels_out=torch.empty(4, 32, 32, 3, dtype=torch.uint8, device=DEV))
    assert pix.shape == (4, 32, 32, 3) and s.plan.n_x0_slots == 2
    s, wrap = presets.dit(ddim_triple(10), batch=2, vae_scale=True)
    dit_net = DiT(dim=64, depth=1, heads=4).reinit_output().to(DEV).eval()
    z = s.sample(wrap(dit_net, torch.tensor([1, 2], device=DEV)))
    assert z.shape == (2, 4, 32, 32) and s.plan.markov and abs(s.final_scale - 1 / 0.18215) < 1e-12 and torch.isfinite(z).all()
    s, wrap = presets.sd3(os.path.join(weights_dir, "sd3_step_28_weight_sharp.csv"), batch=1, latent=(16, 16, 16))
    mm = MMDiT(dim=64, depth=1, heads=4, ctx_dim=8, pooled_dim=8, max_grid=16).to(DEV).half().eval()
    c, p = torch.randn(1, 3, 8, device=DEV).half(), torch.randn(1, 8, device=DEV).half()
    out = s.sample(wrap(mm, c, p, -c, -p))
    assert out.dtype == torch.float16 and s.plan.n_x0_slots == 14 and torch.isfinite(out).all()
