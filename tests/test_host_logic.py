"""CPU: host-side logic (matrix formats, launch plan, generators, sharding) and the C-ABI surface.
No compute entry point is exercised here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import naturaldiffusion_b200 as ni
from naturaldiffusion_b200 import _lib, generators
from naturaldiffusion_b200.coeffs import (CoeffTriple, build_plan, ddim_x0_coeffs, flow_match_sigmas, io_score_vp,
                                           load_weight_csv, markov_ratios, spaced_timesteps)
from naturaldiffusion_b200.sampler import shard_range
from oracle import ni_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI surface
def test_library_builds_loads_and_exports_every_declared_symbol():
    from naturaldiffusion_b200 import build
    build.build()
    L = ni.lib()
    header = open(os.path.join(ROOT, "include", "ni_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+)?[A-Za-z_0-9]+\s*\*?\s*(ni_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(L, name) is not None
    assert L.ni_version() == _lib.NI_ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of NiStepDesc as gcc sees the header == the ctypes mirror"""
    fields = [f[0] for f in _lib.NiStepDesc._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "ni_b200.h"\nint main(){printf("%zu", sizeof(NiStepDesc));' + \
          "".join(f'printf(" %zu", offsetof(NiStepDesc, {f}));' for f in fields) + "return 0;}\n"
    c = tmp_path / "layout.c"
    c.write_text(src)
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == C.sizeof(_lib.NiStepDesc)
    assert vals[1:] == [getattr(_lib.NiStepDesc, f).offset for f in fields]


def test_load_flavour_rule_on_the_baseline_shapes():
    """ni_step_flavour: the per-launch choice between streaming (1) and L2-friendly (0) loads.  Without a device the
    library assumes a 126 MiB L2 (B200).  Expected values are the faster flavour measured on the B200
    (profiles/r01_policy_sweep.txt); the DDPM first-order path is the one shape where the rule is 1.7% off."""
    from naturaldiffusion_b200.ops import StepLaunch
    F32, F16 = _lib.NI_F32, _lib.NI_F16

    def flavour(numel, dt, m, n_terms, keep_x0=True, keep_gen=0):
        L = StepLaunch(numel=numel, per_sample=numel // 64, dtype=dt, has_x0=True, x_in=0x1000, out0=0x2000, out1=0x3000 if m == 2 else 0,
                       a=1.0, b0=1.0, b1=0.5, x0_dst=0x4000 if keep_x0 else 0, c_x0=1.0, terms=[(0x10000 + 16 * i, 0.1) for i in range(n_terms)],
                       gens=[(i + 1, 0.1, 0x5000) for i in range(keep_gen)], x_next=0x6000)
        return L.flavour()

    c2, c3, c4, c5 = 4096 * 3072, 16384 * 3072, 1024 * 4096, 64 * 16 * 128 * 128
    assert flavour(c2, F32, 1, 4) == 1 and flavour(c3, F32, 1, 4) == 1          # 100 / 400 MB written per launch: streams
    assert flavour(c2, F32, 1, 1, keep_x0=False) == 0                            # C2 steps 0 and 9 write x_{k+1} only (50 MB)
    assert flavour(c3, F32, 1, 1, keep_x0=False) == 1
    assert flavour(c5, F16, 2, 1, keep_x0=False) == 0                            # SD3 first-order path: 33 MB written, +11% with NA loads
    assert flavour(c5, F16, 2, 14) == 0 and flavour(c5, F16, 2, 5) == 0          # sharp / dense SD3 tables at B 64: 67 MB written
    assert flavour(4 * c5, F16, 2, 10) == 1                                      # B 256: 268 MB written
    assert flavour(c4, F32, 2, 200, keep_gen=1) == 1                             # dense DDPM-250 row: 3 of ~200 tensors written
    assert flavour(c4, F32, 2, 0, keep_x0=False) == 0
    try:
        _lib.set_option("load_policy", 2)
        assert flavour(c5, F16, 2, 1, keep_x0=False) == 1
        _lib.set_option("load_policy", 1)
        assert flavour(c3, F32, 1, 4) == 0
    finally:
        _lib.set_option("load_policy", 0)
    assert ni.lib().ni_step_flavour(None) == -1


def test_argument_validation_needs_no_gpu():
    L = ni.lib()
    assert L.ni_step(None, None) == -1 and b"NULL" in L.ni_last_error()
    d = _lib.NiStepDesc()
    d.numel, d.per_sample = 10, 3
    assert L.ni_step(C.byref(d), None) == -1 and b"multiple" in L.ni_last_error()
    d.numel, d.per_sample, d.n_terms = 12, 3, 10_000
    assert L.ni_step(C.byref(d), None) == -3
    assert L.ni_weighted_sum(None, None, 513, None, 0, 0, 0, 1.0, None) == -3
    assert L.ni_philox_normal(None, 4, 0, 0, 0, 0, None) == -1
    assert L.ni_weighted_sum(None, None, 0, None, 0, 0, 0, 1.0, None) == 0  # empty input: nothing to do


def test_missing_library_is_loud(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libni_b200.so")
    with pytest.raises(ni.NiError, match="no CPU fallback"):
        _lib.lib()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "naturaldiffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


# ------------------------------------------------------------------ matrices / plan
def test_liveness_matches_survey(weights_dir):
    want = {"step_5_weight_00": 2, "step_10_weight_42": 4, "step_15_weight_173": 5}
    for name, slots in want.items():
        p = build_plan(CoeffTriple.from_npz(os.path.join(weights_dir, name + ".npz")))
        assert p.n_x0_slots == slots and p.n_eps_slots == 0 and p.eps0_last_use == p.K - 1
    sig = flow_match_sigmas(28)
    assert build_plan(CoeffTriple.from_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight_sharp.csv"), sig)).n_x0_slots == 14
    assert build_plan(CoeffTriple.from_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight.csv"), sig)).n_x0_slots == 27


def test_plan_never_reads_a_recycled_slot(weights_dir):
    """simulate the ring: every (row, column) read must find that column still resident in its slot"""
    triples = [CoeffTriple.from_npz(os.path.join(weights_dir, n + ".npz")) for n in ("step_10_weight_42", "step_15_weight_173")]
    triples += [generators.ddpm_triple(24), CoeffTriple.from_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight_sharp.csv"))]
    for t in triples:
        p = build_plan(t)
        x0_res, eps_res = {}, {}
        for k, s in enumerate(p.steps):
            for j, _ in s.hist:
                assert x0_res.get(p.x0_slot_of[j]) == j, (t.name, k, j)
            for j, _ in s.eps:
                if j > 0:
                    assert eps_res.get(p.eps_slot_of[j]) == j
            if s.keep_x0:
                x0_res[s.x0_slot] = k
            if s.keep_fresh:
                eps_res[s.fresh_slot] = k + 1
        # rebuilt row == matrix row
        for k, s in enumerate(p.steps):
            row = np.zeros(t.K)
            for j, c in s.hist:
                row[j] = c
            row[k] = s.c_x0
            assert np.array_equal(row, np.where(t.A[k] != 0, t.A[k], 0.0))


def test_ring_plan_on_random_sparse_matrices():
    """liveness / slot allocation under arbitrary zero patterns: simulate the ring on the host and check that every read
    finds its column resident, that slots are never shared by two live columns, and that the slot count equals the
    peak number of simultaneously live columns (the allocation is optimal for interval graphs)"""
    for seed in range(60):
        rng = np.random.default_rng(seed)
        K = int(rng.integers(1, 25))
        A = np.tril(rng.standard_normal((K, K))) * (rng.random((K, K)) < rng.uniform(0.05, 1.0))
        A[np.arange(K), np.arange(K)] = 1.0
        B = np.zeros((K, K + 1))
        B[:, 0] = rng.standard_normal(K) * (rng.random(K) < 0.7)
        if seed % 2:
            for k in range(K):
                B[k, 1:k + 2] = rng.standard_normal(k + 1) * (rng.random(k + 1) < rng.uniform(0.1, 0.9))
        t = CoeffTriple(A, B, np.zeros((K + 1, 3)))
        p = build_plan(t)
        x0_res, eps_res, peak_x0, peak_eps = {}, {}, 0, 0
        for k, s in enumerate(p.steps):
            for j, c in s.hist:
                assert x0_res.get(p.x0_slot_of[j]) == j and c == A[k, j]
            for j, c in s.eps:
                assert c == B[k, j]
                if j > 0:
                    assert eps_res.get(p.eps_slot_of[j]) == j
            if s.keep_x0:
                assert x0_res.get(s.x0_slot) is None or max(kk for kk in range(K) if A[kk, x0_res[s.x0_slot]] != 0) < k
                x0_res[s.x0_slot] = k
            if s.keep_fresh:
                eps_res[s.fresh_slot] = k + 1
            # a slot is busy at step k if its column is still READ at step k or later (no slot is recycled inside the
            # kernel that reads it), plus the column produced at step k if a later row needs it
            live_x0 = sum(1 for j in range(k) if any(A[kk, j] != 0 for kk in range(k, K))) + int(any(A[kk, k] != 0 for kk in range(k + 1, K)))
            live_eps = sum(1 for j in range(1, k + 1) if any(B[kk, j] != 0 for kk in range(k, K))) + int(any(B[kk, k + 1] != 0 for kk in range(k + 1, K)))
            peak_x0, peak_eps = max(peak_x0, live_x0), max(peak_eps, live_eps)
            assert {j for j, _ in s.hist} == {j for j in range(k) if A[k, j] != 0}
            assert {j for j, _ in s.eps} == {j for j in range(k + 1) if B[k, j] != 0}
            assert (s.fresh is not None and s.fresh != 0.0) == (B[k, k + 1] != 0)
        assert p.n_x0_slots == peak_x0 and p.n_eps_slots == peak_eps, (seed, p.n_x0_slots, peak_x0, p.n_eps_slots, peak_eps)


def test_algorithmic_units_formula(weights_dir):
    """bytes(k) = s*N*[m + 1 + write_x0 + (nnzA-1) + nnzB_stored + w_eps + 1]"""
    t = CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz"))
    p = build_plan(t)
    nnz = (t.A != 0).sum(1)
    for k in range(10):
        keep = 1 if p.steps[k].keep_x0 else 0
        assert p.units(k, 1) == 1 + 1 + keep + (nnz[k] - 1) + 1 + 0 + 1
        assert p.units(k, 1, eps0_stored=False) == p.units(k, 1) - 1
    assert p.total_units(1) == 65  # SURVEY's 67 minus the two x0 writes no later row reads (columns 0 and 9)
    # dense stochastic rows (C4, ddpm_250), counted independently from the matrix: the last DDPM row has
    # coeff_xt = 0, so it reads no history and x0_248 / eps_249 / x0_249 / eps_250 are never kept
    t = generators.ddpm_triple(250)
    d = build_plan(t)
    nzA, nzB = t.A != 0, t.B != 0
    want = 0
    for k in range(250):
        keep_x0 = nzA[k + 1:, k].any()
        keep_eps = nzB[k + 1:, k + 1].any()
        want += 2 + 1 + int(keep_x0) + (nzA[k, :k].sum()) + nzB[k, :k + 1].sum() + int(keep_eps) + 1
    assert d.total_units(2) == want == 63497


def test_npz_roundtrip_and_shapes(tmp_path, weights_dir):
    t = CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz"))
    assert t.B.shape == (10, 11)  # K x K file normalised to K x (K+1)
    t.save_npz(tmp_path / "x.npz")
    A, B, node = O.load_triple(tmp_path / "x.npz")  # by position, like the reference
    assert np.array_equal(A, t.A) and np.array_equal(B, t.B) and np.array_equal(node, t.node)
    with pytest.raises(ValueError):
        CoeffTriple(np.triu(np.ones((3, 3))), np.zeros((3, 4)), np.zeros((4, 3)))
    assert np.array_equal(load_weight_csv(os.path.join(weights_dir, "sd3_step_28_weight.csv")),
                          O.load_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight.csv")))
    assert np.array_equal(flow_match_sigmas(28), O.sd3_sigmas(28))


def test_companion_csv_writer_reproduces_the_reference_files(golden_dir, tmp_path):
    """`save_coeff_matrix` (src/Utils.py:30-45) writes a rounded csv next to every npz; CoeffTriple.save_companion_csv
    reproduces the shipped files byte for byte (up to the CRLF of the Windows box they were written on) -- discrete
    (%03d) and continuous (%0.3f) time labels, -0.0 entries, the sum column.  All 44 files when the reference tree is here."""
    import glob
    m = np.load(os.path.join(golden_dir, "reference_matrices.npz"))
    cases = [(os.path.join(golden_dir, "reference_csv", os.path.basename(k) + ".csv"), m[k + "/A"], m[k + "/B"], m[k + "/node"])
             for k in ("ddim/ddim_018", "ddpm/ddpm_sympy_018", "dpmsolverpp/dpmsolverpp2s_018", "deis/deis_tab_100")]
    for npz in sorted(glob.glob("/root/reference/results/*/*.npz")):
        if os.path.isfile(npz[:-4] + ".csv"):
            cases.append((npz[:-4] + ".csv",) + tuple(np.load(npz).values()))
    assert len(cases) >= 4
    for csv_path, A, B, node in cases:
        out = tmp_path / "t.csv"
        CoeffTriple(A, B, node).save_companion_csv(out)
        assert out.read_bytes() == open(csv_path, "rb").read().replace(b"\r\n", b"\n"), csv_path


def test_sd3_weight_table_writer(weights_dir, tmp_path):
    """save_weight_csv writes what src/SD3NaturalInference.py:196 reads: the generated default table reproduces the
    shipped csv byte for byte (up to CRLF); the hand-edited sharp table (integer cells) round-trips by value."""
    from naturaldiffusion_b200.coeffs import save_weight_csv
    sig = flow_match_sigmas(28)
    for name, exact in (("sd3_step_28_weight.csv", True), ("sd3_step_28_weight_sharp.csv", False)):
        src = os.path.join(weights_dir, name)
        W = load_weight_csv(src)
        out = tmp_path / name
        save_weight_csv(W, sig, out)
        assert np.array_equal(load_weight_csv(out), W)
        if exact:
            assert out.read_bytes() == open(src, "rb").read().replace(b"\r\n", b"\n")
    # the default table IS round(100 * (sigma_i - sigma_{i+1}), 2) on every row
    W = load_weight_csv(os.path.join(weights_dir, "sd3_step_28_weight.csv"))
    from naturaldiffusion_b200.coeffs import flow_euler_weight_table
    assert np.array_equal(W, flow_euler_weight_table(sig))
    with pytest.raises(ValueError):
        save_weight_csv(W[:5], sig, tmp_path / "bad.csv")


def test_optimised_matrices_rebuild_from_relative_patterns(weights_dir):
    """weights/step_{10,15}_weight_*.npz are 2-decimal relative patterns scaled to rowsum = alpha, B[:,0] = sigma:
    extracting the patterns and rebuilding gives the shipped A to 1 ulp; step_5 needs its unrounded ratios.  The
    quadratic VP node grid matches the stored one to float32 round-off."""
    for name, dec in (("step_10_weight_42", 2), ("step_15_weight_173", 2), ("step_5_weight_00", None)):
        t = CoeffTriple.from_npz(os.path.join(weights_dir, name + ".npz"))
        pats = generators.relative_patterns(t, decimals=dec)
        r = generators.relative_pattern_triple(pats, t.node)
        assert np.abs(r.A - t.A).max() < 5e-16 and np.array_equal(r.B, t.B)
        band = [p[np.flatnonzero(p)[0]:] for p in pats]          # right-aligned band form
        assert np.abs(generators.relative_pattern_triple(band, t.node).A - r.A).max() < 5e-16
        node = generators.vp_quadratic_node(t.K)
        assert np.abs(node[:, 0] - t.node[:, 0]).max() < 2e-7 and np.abs(node[:, 1:] - t.node[:, 1:]).max() < 2e-6
        assert build_plan(r).n_x0_slots == build_plan(t).n_x0_slots
    assert list(generators.relative_patterns(CoeffTriple.from_npz(os.path.join(weights_dir, "step_15_weight_173.npz")))[5]) == [0, 0.30, -0.14, 0.56, -0.77, 1]
    with pytest.raises(ValueError):
        generators.relative_pattern_triple([[1.0], [1.0, -1.0]], generators.vp_quadratic_node(2))


def test_generators_match_reference_matrices(golden_dir):
    m = np.load(os.path.join(golden_dir, "reference_matrices.npz"))
    for fam, fn in (("ddim", generators.ddim_triple), ("ddpm", generators.ddpm_triple)):
        for K in (18, 24, 100):
            t = fn(K)
            key = f"{fam}/{fam}_{K:03d}"
            assert np.abs(t.A - m[key + "/A"]).max() < 1e-14 and np.abs(t.B - m[key + "/B"]).max() < 1e-14
            assert np.array_equal(t.node, m[key + "/node"])
            assert markov_ratios(t) is not None
    for K in (18, 24):
        t = generators.flow_euler_triple(K)
        key = f"flow_euler/flow_euler_simpy_{K:03d}"
        # the reference grid is ascending sigma with the matrix in sampling order
        assert np.abs(t.A - m[key + "/A"]).max() < 1e-14 and np.abs(t.B - m[key + "/B"]).max() < 1e-14
    for K in (10, 250):  # the two matrices BASELINE's configs need
        for fn, ofn in ((generators.ddim_triple, O.ddim_triple), (generators.ddpm_triple, O.ddpm_triple)):
            t, (A, B, node) = fn(K), ofn(K)
            assert np.abs(t.A - A).max() < 1e-14 and np.abs(t.B - B).max() < 1e-14 and np.array_equal(t.node, node)
    m42 = CoeffTriple(*[m["dpmsolverpp/dpmsolverpp2s_024/" + n] for n in ("A", "B", "node")])
    assert markov_ratios(m42) is None  # higher-order solvers are not Markov in the rows


def test_markov_structure_detection_and_identity(weights_dir):
    """first-order x0 structure: row_k . history == c_k * x_k + R_k . eps, for every family that has it"""
    from naturaldiffusion_b200.coeffs import markov_ratios
    fams = {"ddpm24": generators.ddpm_triple(24), "ddim100": generators.ddim_triple(100), "flow18": generators.flow_euler_triple(18),
            "sd3": CoeffTriple.from_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight.csv"))}
    for name, t in fams.items():
        cs, R = markov_ratios(t)
        K = t.K
        prev = np.zeros(2 * K + 1)
        prev[K] = 1.0  # x_0 = eps_0
        for k in range(K):
            row = np.concatenate([t.A[k], t.B[k]])
            rec = cs[k] * prev
            rec[k] = t.A[k, k]
            rec[K:K + k + 1] += R[k, :k + 1]
            rec[K + k + 1] = t.B[k, k + 1]
            assert np.abs(rec - row).max() < 1e-11, (name, k)
            prev = row
        p = build_plan(t, markov=True)
        assert p.markov and p.n_x0_slots == 0
        assert (R != 0).sum() == (27 if name == "sd3" else 0)  # only the 2-decimal SD3 table needs an eps_0 correction
        assert p.total_units(2) == sum(2 + 1 + 1 + len(s.eps) for s in p.steps)
    for name in ("step_10_weight_42", "step_15_weight_173"):
        assert markov_ratios(CoeffTriple.from_npz(os.path.join(weights_dir, name + ".npz"))) is None
    assert markov_ratios(CoeffTriple.from_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight_sharp.csv"))) is None
    with pytest.raises(ValueError):
        build_plan(CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz")), markov=True)


def test_coefficient_space_tracer_reproduces_reference_dpm_solver_matrices(golden_dir, weights_dir):
    """f2: running a linear sampler on coefficient-space vectors yields its matrix; DPM-Solver-2S and DPM-Solver++(2S)
    reproduce the 8 matrices the reference derived with sympy (results/dpmsolver*/), and the multistep ++(2M) matrix on
    the quadratic grid has the nodes of weights/step_15_weight_173 (BASELINE config 3)"""
    m = np.load(os.path.join(golden_dir, "reference_matrices.npz"))
    for fam, fn in (("dpmsolverpp/dpmsolverpp2s", generators.dpm_solver_pp_2s_triple), ("dpmsolver/dpmsolver2s", generators.dpm_solver_2s_triple)):
        for K in (18, 24, 100, 200):
            t = fn(K // 2)
            key = f"{fam}_{K:03d}"
            assert np.abs(t.A - m[key + "/A"]).max() < 1e-12 and np.abs(t.B - m[key + "/B"]).max() < 1e-12, key
            assert np.abs(t.node - m[key + "/node"]).max() < 1e-12
    t = generators.dpm_solver_pp_2m_triple(15)
    ref = CoeffTriple.from_npz(os.path.join(weights_dir, "step_15_weight_173.npz"))
    assert np.abs(t.node - ref.node).max() < 1e-6  # same quadratic grid / VP marginals (the file carries fp32 round-off)
    assert np.abs(t.A.sum(1) - t.node[1:, 1]).max() < 7e-3 and np.count_nonzero(t.B[:, 1:]) == 0
    assert build_plan(t).n_x0_slots == 15 - 1  # a true multistep solver is dense: every x0 stays live
    from naturaldiffusion_b200.coeffs import markov_ratios
    assert markov_ratios(t) is None
    # the tracer on a first-order sampler gives the closed form
    ns, ts = generators.VPLinearSchedule(), np.linspace(1.0, 1e-3, 11)
    tr = generators.CoefficientTracer(10, ns)
    x = tr.noise()
    for i in range(10):
        y = tr.model_x0(x, ts[i])
        x = ns.sigma(ts[i + 1]) / ns.sigma(ts[i]) * x + (ns.alpha(ts[i + 1]) - ns.sigma(ts[i + 1]) / ns.sigma(ts[i]) * ns.alpha(ts[i])) * y  # DDIM
    d = tr.finish(x, ts[-1])
    assert markov_ratios(d) is not None and np.abs(d.A.sum(1) + d.B[:, 0] * 0 - d.node[1:, 1]).max() < 7e-3


def test_deis_generator_matches_reference_matrices(golden_dir):
    """DEIS tAB3 on the quadratic grid (the reference needs jax for it, src/AnalyzeDEIS.py): float64 numpy restatement of
    the same Riemann-sum coefficients reproduces results/deis/deis_tab_{100,200} to the jax-float32 noise of the files"""
    m = np.load(os.path.join(golden_dir, "reference_matrices.npz"))
    for K in (100, 200):
        t = generators.deis_tab_triple(K)
        key = f"deis/deis_tab_{K:03d}"
        assert np.abs(t.A - m[key + "/A"]).max() < 1e-5 and np.abs(t.B - m[key + "/B"]).max() < 1e-5
        assert np.abs(t.node - m[key + "/node"]).max() < 1e-6
    c = O.deis_tab_coefficients(generators.quadratic_time_grid(15))
    assert c.shape == (15, 5) and np.all(c[0, 2:] == 0) and np.all(c[1, 3:] == 0) and c[5, 4] != 0  # order ramp 0,1,2,3


def test_every_shipped_matrix_is_reproduced(golden_dir):
    """all 44 coefficient matrices under the reference's results/ (every sampler family it analyses: DDPM, DDIM, VP
    ODE/SDE Euler, Heun, DPM-Solver-2S/3S, DPM-Solver++(2S/3S), DEIS tAB3, flow Euler) come out of generators.py"""
    m = np.load(os.path.join(golden_dir, "reference_matrices.npz"))
    fam = {
        "ddim/ddim": lambda K: generators.ddim_triple(K), "ddpm/ddpm": lambda K: generators.ddpm_triple(K),
        "ddpm/ddpm_sympy": lambda K: generators.ddpm_triple(K), "deis/deis_tab": lambda K: generators.deis_tab_triple(K),
        "dpmsolver/dpmsolver2s": lambda K: generators.dpm_solver_2s_triple(K // 2),
        "dpmsolver/dpmsolver3s": lambda K: generators.dpm_solver_3s_triple(K // 3),
        "dpmsolverpp/dpmsolverpp2s": lambda K: generators.dpm_solver_pp_2s_triple(K // 2),
        "dpmsolverpp/dpmsolverpp3s": lambda K: generators.dpm_solver_3s_triple(K // 3, plus_plus=True),
        "euler_heun/ode_euler": lambda K: generators.vp_euler_triple(K, "ode"),
        "euler_heun/sde_euler": lambda K: generators.vp_euler_triple(K, "sde"),
        "euler_heun/ode_heun": lambda K: generators.vp_euler_triple(K // 2, "heun"),
        "flow_euler/flow_euler_simpy": lambda K: generators.flow_euler_triple(K),
    }
    keys = sorted({k.rsplit("/", 1)[0] for k in m.files if k.endswith("/A")})
    done = 0
    for key in keys:
        name, K = key.rsplit("_", 1)
        t = fam[name](int(K))
        tol = 1e-5 if "deis" in key else 1e-12  # the DEIS files carry jax-float32 quadrature noise
        assert np.abs(t.A - m[key + "/A"]).max() < tol and np.abs(t.B - m[key + "/B"]).max() < tol, key
        node_ref = m[key + "/node"].copy()
        if "sympy" in key:  # the sympy variant starts from alpha(T) = 0.0064 instead of 0 and lists the marginals of the grid itself
            assert np.array_equal(t.node[1:-1, 0], node_ref[1:-1, 0]) and np.abs(t.node[1:-1] - node_ref[1:-1]).max() < 1e-4, key
        else:
            assert np.abs(t.node - node_ref).max() < max(tol, 1e-6), key
        done += 1
    for famname, fn in (("ddim", generators.ddim_triple), ("ddpm", generators.ddpm_triple)):  # the two K=500 files, by digest
        t = fn(500)
        d = np.concatenate([t.A.sum(0), t.A.sum(1), t.B.sum(0), t.B.sum(1), np.diag(t.A), t.node.ravel()])
        assert np.abs(d - m[f"{famname}/{famname}_500/digest"]).max() < 1e-11
        done += 1
    assert done == 44


def test_config_matrices_match_the_reference_generators(golden_dir):
    """ddim_010 (BASELINE config C1) and ddpm_250 (C4) are not shipped by the reference; tests/golden/config_matrices.npz
    holds what the reference's OWN generators output for them (make_golden.py::gen_config_matrices).  Both the product
    generators and the oracle's restatement reproduce them."""
    z = np.load(os.path.join(golden_dir, "config_matrices.npz"))
    for key, fn, ofn, K in (("ddim_010", generators.ddim_triple, O.ddim_triple, 10), ("ddpm_250", generators.ddpm_triple, O.ddpm_triple, 250),
                            ("ddpm_010", generators.ddpm_triple, O.ddpm_triple, 10)):
        t, (A, B, node) = fn(K), ofn(K)
        for got_A, got_B, got_node in ((t.A, t.B, t.node), (A, B, node)):
            assert np.abs(got_A - z[key + "/A"]).max() < 1e-15 and np.abs(got_B - z[key + "/B"]).max() < 1e-15
            assert np.array_equal(got_node, z[key + "/node"])


def test_schedule_helpers():
    assert spaced_timesteps(1000, 10) == [0, 111, 222, 333, 444, 555, 666, 777, 888, 999]
    assert spaced_timesteps(1000, 10) == O.spaced_steps(1000, 10)
    c1, c2, idx = ddim_x0_coeffs(10)
    assert abs(c1[0] - 157.41046) < 1e-4 and abs(c2[-1] - 0.01) < 1e-4 and idx[0] == 999
    t = O.skip_tables(24)
    assert np.allclose(c1 if len(c1) == 24 else ddim_x0_coeffs(24)[0], t["xt2x0"][::-1])
    io = io_score_vp(np.array([[1.0, 0.0066, 1.0], [0.5, 0.5, 0.8]]))
    assert abs(io[0][0] - 1 / 0.0066) < 1e-9 and io[0][1] < 0 and io[0][2] == 0.0


def test_shard_range_partitions():
    for total, world in ((4096, 8), (50000, 8), (7, 3), (5, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_dropins_fit_the_reference_modules():
    """(build container only) the reference scripts define exactly the names `dropin.install` replaces, with call
    signatures the replacements accept.  The reference sources are PARSED (ast), not executed."""
    import inspect
    import types
    from oracle import ref_loader
    if not ref_loader.present():
        pytest.skip("reference tree not present (GPU box)")
    from naturaldiffusion_b200 import dropin
    v = ref_loader.signatures("src/ValidateNaturalInference.py")
    s3 = ref_loader.signatures("src/SD3NaturalInference.py")
    cf = ref_loader.signatures("src/CIFAR10NaturalInference.py")
    assert v["weighted_sum"] == ["weights", "seq_elem"]
    assert cf["weighted_sum"] == ["past_x0_coeff", "seq_x0"]
    assert s3["weighted_sum"] == ["seq_xstarts", "weights"]
    assert s3["euler_weighted_sum"] == ["seq_xstarts", "cliplen"]
    assert cf["data_fn"] == list(inspect.signature(dropin.data_fn).parameters)
    assert len(inspect.signature(dropin.weighted_sum).parameters) == 2
    assert list(inspect.signature(dropin.euler_weighted_sum).parameters) == ["seq_xstarts", "cliplen"]
    # install() patches exactly the hot-path names a module defines
    mv = types.SimpleNamespace(**{k: None for k in v})
    ms = types.SimpleNamespace(**{k: None for k in s3})
    mc = types.SimpleNamespace(**{k: None for k in cf})
    assert set(dropin.install(mv)) == {"weighted_sum"} and mv.weighted_sum is dropin.weighted_sum
    assert set(dropin.install(ms)) == {"weighted_sum", "euler_weighted_sum"}
    assert set(dropin.install(mc)) == {"weighted_sum", "data_fn"}


def test_kxk_epsilon_matrix_uses_column_zero_only():
    """weights/*.npz carry a K x K past_epsilon_coeff whose reader uses column 0 only (src/CIFAR10NaturalInference.py:303)"""
    import warnings
    K = 4
    A = np.tril(np.ones((K, K)))
    B = np.zeros((K, K)); B[:, 0] = 0.5; B[2, 1] = 0.25
    node = np.stack([np.linspace(1, 0, K + 1)] * 3, axis=1)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        t = CoeffTriple(A, B, node)
    assert w and "column 0" in str(w[0].message)
    assert t.B.shape == (K, K + 1) and np.all(t.B[:, 0] == 0.5) and np.all(t.B[:, 1:] == 0)
    assert all(s.fresh is None for s in build_plan(t).steps)


def test_bind_rank_cpus_gives_disjoint_slices(monkeypatch):
    """8 ranks whose GPUs hang off the same NUMA node (this pool's VMs) share its CPUs out in disjoint, equal slices; ranks on
    different nodes keep their own node"""
    from naturaldiffusion_b200 import hostutil
    bound = {}
    monkeypatch.setattr(hostutil.os, "sched_setaffinity", lambda pid, cpus: bound.__setitem__("cpus", set(cpus)))
    monkeypatch.setattr(hostutil, "gpu_local_cpus", lambda i: set(range(32)))
    slices = []
    for r in range(8):
        assert hostutil.bind_rank_cpus(r, 8) == 4
        slices.append(bound["cpus"])
    assert set().union(*slices) == set(range(32)) and sum(len(x) for x in slices) == 32
    monkeypatch.setattr(hostutil, "gpu_local_cpus", lambda i: set(range(16)) if i < 4 else set(range(16, 32)))
    assert hostutil.bind_rank_cpus(5, 8) == 4 and bound["cpus"] == {20, 21, 22, 23}
    assert hostutil.bind_rank_cpus(0, 1) == 16
    monkeypatch.setattr(hostutil, "gpu_local_cpus", lambda i: None)
    assert hostutil.bind_rank_cpus(0, 8) is None


def test_multiply_shift_sample_index_formula():
    """the specialised step kernels get the sample index of a CTA as umulhi(blockIdx, mul) >> shr (csrc/ni_step_lean.cuh
    `fast_divisor`); the same arithmetic in Python is exact for every block index below 2^31"""
    rng = np.random.default_rng(0)

    def fast_divisor(d):
        if d <= 1:
            return 0, 0
        lg = int(np.ceil(np.log2(d)))
        if (1 << lg) < d:
            lg += 1
        p = 31 + lg
        return ((1 << p) + d - 1) // d, p - 32

    for d in list(range(1, 300)) + [384, 768, 1000, 4097, 65535, 65536, 1 << 20, (1 << 24) - 1]:
        mul, shr = fast_divisor(d)
        assert mul < (1 << 32)
        ns = np.concatenate([np.arange(0, 2000), rng.integers(0, 1 << 31, 2000), [(1 << 31) - 1, d - 1, d, 2 * d - 1]]).astype(np.uint64)
        q = ns if mul == 0 else ((ns * np.uint64(mul)) >> np.uint64(32)) >> np.uint64(shr)
        assert np.array_equal(q, ns // np.uint64(d)), d


def test_host_relay_pairing_from_probe_times():
    """hostutil.relay_pairs: the measured 8-GPU box (profiles/r02_host_copy_probe_n8.json: GPUs 0-3 at ~12.2 GB/s, GPUs 4-7 at
    ~19 GB/s with all ranks copying) pairs each far-socket rank with a near one; uniform boxes keep the direct route"""
    from naturaldiffusion_b200.hostutil import relay_pairs
    gbs = [12.15, 12.18, 12.14, 12.16, 19.2, 19.04, 18.94, 18.9]
    ms = [100.0 / g for g in gbs]
    assert relay_pairs(ms) == {0: 4, 1: 5, 2: 6, 3: 7}
    assert relay_pairs(ms[::-1]) == {4: 0, 5: 1, 6: 2, 7: 3}
    assert relay_pairs([1.0] * 8) is None and relay_pairs([1.0, 1.1]) is None and relay_pairs([1.0]) is None
    assert relay_pairs([1.0, 1.0, 1.0, 2.0]) is None          # one straggler is not a slower HALF
    assert relay_pairs([1.0, 2.0, 1.0]) is None                # odd world
    assert relay_pairs([1.0, 2.0]) == {1: 0}
    assert relay_pairs([1.0, 1.0], "force") == {0: 1, 1: 0}


def test_choose_host_relay_without_a_group_keeps_the_direct_route():
    """world 1, mode "off", or no initialised process group: no probe, no CUDA call, direct route"""
    from naturaldiffusion_b200.hostutil import choose_host_relay
    for kw in (dict(rank=0, world=1, device="cuda:0"), dict(rank=0, world=8, device="cuda:0", mode="off"), dict(rank=1, world=2, device="cuda:1")):
        peers, info = choose_host_relay(**kw)
        assert peers == {"d2h": None, "bidir": None} and "mode" in info


def test_flow_euler_table_with_a_clip_window_matches_euler_weighted_sum():
    """coeffs.flow_euler_weight_table(cliplen=c, rounded=False) -> CoeffTriple.from_sd3_table gives, row by row, the update of
    src/SD3NaturalInference.py:129 with `euler_weighted_sum(seq_xstarts, cliplen)` (:61-69, the `[-cliplen:]` window):
    x_{k+1} = sigma_{k+1} noise + (1 - sigma_{k+1}) * sum_window w_j x0_j / sum_window w_j  -- checked against the oracle's
    restatement of that function on random tensors."""
    import torch
    from naturaldiffusion_b200.coeffs import CoeffTriple, flow_euler_weight_table, flow_match_sigmas
    from oracle import ni_oracle as O
    sig = flow_match_sigmas(12).astype(np.float64)
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(2, 3, 4, 4, generator=g, dtype=torch.float64) for _ in range(12)]
    for clip in (0, 1, 3, 12, 40):
        W = flow_euler_weight_table(sig, cliplen=clip, rounded=False)
        for k in range(12):
            lo = 0 if clip == 0 else max(0, k + 1 - clip)
            assert np.all(W[k, :lo] == 0) and np.all(W[k, lo:k + 1] != 0) and np.all(W[k, k + 1:] == 0)
        t = CoeffTriple.from_sd3_table(W, sig)
        seq = []
        for k in range(12):
            seq.append([sig[k] - sig[k + 1], xs[k]])
            ref = (1.0 - sig[k + 1]) * O.euler_weighted_sum(seq, clip)[1]
            got = sum(t.A[k, j] * xs[j] for j in range(k + 1))
            assert (got - ref).abs().max().item() < 1e-13 and t.B[k, 0] == sig[k + 1]
    # the rounded, unclipped table is still the shipped csv's rule
    assert np.array_equal(flow_euler_weight_table(sig), flow_euler_weight_table(sig, 0, True))
