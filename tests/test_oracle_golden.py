"""CPU: pin the oracle (oracle/) against the golden vectors produced by the reference's own functions
(tests/golden/make_golden.py) and against the published Philox known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import ni_oracle as O
from oracle import philox
from toy_models import ToyEps

torch.set_grad_enabled(False)


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ------------------------------------------------------------------ Philox
KAT = [  # Random123 v1.09 kat_vectors, philox4x32-10
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(ctr, key, want):
    assert philox.philox4x32_10(ctr, key) == want


def test_philox_normal_moments_and_offsets():
    a = philox.normal((1 << 20,), seed=888, tensor_id=0)
    assert abs(a.mean()) < 4e-3 and abs(a.std() - 1) < 4e-3
    assert abs(((a - a.mean()) ** 4).mean() / a.var() ** 2 - 3) < 0.03
    # keyed by global element index: any shard offset reproduces the same values
    for off in (0, 4, 5, 1023, 4096):
        b = philox.normal((1000,), seed=888, tensor_id=0, elem_offset=off)
        assert np.array_equal(b, a[off:off + 1000])
    c = philox.normal((4096,), seed=888, tensor_id=1)
    assert abs(np.corrcoef(a[:4096], c)[0, 1]) < 0.06
    assert not np.array_equal(philox.normal((64,), seed=889, tensor_id=0), a[:64])


# ------------------------------------------------------------------ matrices and schedules
def test_generators_match_shipped_matrices(golden_dir):
    m = _g(golden_dir, "reference_matrices.npz")
    checked = 0
    for fam, fn in (("ddim", O.ddim_triple), ("ddpm", O.ddpm_triple)):
        for K in (18, 24, 100):
            A, B, node = fn(K)
            key = f"{fam}/{fam}_{K:03d}"
            assert np.abs(A - m[key + "/A"]).max() < 1e-15
            assert np.abs(B - m[key + "/B"]).max() < 1e-15
            assert np.array_equal(node, m[key + "/node"])
            checked += 1
        A, B, node = fn(500)
        d = np.concatenate([A.sum(0), A.sum(1), B.sum(0), B.sum(1), np.diag(A), node.ravel()])
        assert np.abs(d - m[f"{fam}/{fam}_500/digest"]).max() < 1e-12
    for K in (18, 24):
        A, B, node = O.flow_euler_triple(K)
        key = f"flow_euler/flow_euler_simpy_{K:03d}"
        assert np.abs(A - m[key + "/A"]).max() < 1e-15 and np.abs(B - m[key + "/B"]).max() < 1e-15
        assert np.abs(node - m[key + "/node"]).max() < 1e-15
    # closed form vs the reference's sympy expansion: identical except node[0,1]
    for K in (18, 24, 100):
        A, B, node = O.ddpm_triple(K)
        key = f"ddpm/ddpm_sympy_{K:03d}"
        assert np.abs(A - m[key + "/A"]).max() < 1e-14 and np.abs(B - m[key + "/B"]).max() < 1e-14
    assert checked == 6


def test_matrix_invariants(golden_dir):
    """rowsum(A) ~ alpha, ||B row|| ~ sigma for every shipped matrix family that satisfies them."""
    m = _g(golden_dir, "reference_matrices.npz")
    keys = sorted({k.rsplit("/", 1)[0] for k in m.files if k.endswith("/A")})
    assert len(keys) >= 40
    for key in keys:
        A, B, node = m[key + "/A"], m[key + "/B"], m[key + "/node"]
        K = A.shape[0]
        assert A.shape == (K, K) and B.shape == (K, K + 1) and node.shape == (K + 1, 3)
        assert np.all(np.triu(A, 1) == 0) and np.all(np.triu(B, 2) == 0)
        if "heun" in key:
            continue  # reference quirk: 2nd Heun stage uses y_coeff_s (appendix D.6)
        assert np.abs(A.sum(1) - node[1:, 1]).max() < 7e-3, key
        assert np.abs(np.linalg.norm(B, axis=1) - node[1:, 2]).max() < 2e-2, key


def test_sd3_sigmas_match_csv_header(weights_dir):
    sig = O.sd3_sigmas().astype(np.float64)
    for name in ("sd3_step_28_weight.csv", "sd3_step_28_weight_sharp.csv"):
        with open(os.path.join(weights_dir, name)) as f:
            header = [float(v) for v in f.readline().strip().split(",")[1:]]
        assert np.abs(np.round(sig[1:], 2) - np.array(header)).max() < 1e-9
    W = O.load_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight.csv"))
    assert np.abs(W[27] - np.round(100 * (sig[:-1] - sig[1:]), 2)).max() < 1e-9  # csv body = 100*(sigma_j - sigma_{j+1})


def test_weights_files(weights_dir):
    for name, K, nnz in (("step_5_weight_00", 5, None), ("step_10_weight_42", 10, 27), ("step_15_weight_173", 15, 61)):
        A, B, node = O.load_triple(os.path.join(weights_dir, name + ".npz"))
        assert A.shape == (K, K) and B.shape == (K, K) and node.shape == (K + 1, 3)
        assert np.abs(A.sum(1) - node[1:, 1]).max() < 1e-12 and np.abs(B[:, 0] - node[1:, 2]).max() < 1e-12
        if nnz:
            assert int((A != 0).sum()) == nnz


# ------------------------------------------------------------------ loops
@pytest.mark.parametrize("name", ["step_5_weight_00", "step_10_weight_42", "step_15_weight_173"])
def test_cifar_loop_matches_reference(golden_dir, weights_dir, name):
    g = _g(golden_dir, "cifar_loop.npz")
    A, B, node = O.load_triple(os.path.join(weights_dir, name + ".npz"))
    net = ToyEps(3, seed=11)
    score_fn = O.make_vp_score_fn(lambda x, labels: net(x, labels))
    x, trace = O.cifar_ni_loop(A, B, node, score_fn, torch.from_numpy(g[name + "/noise"]))
    for k, tr in enumerate(trace):
        assert np.array_equal(tr["x_next"].numpy(), g[name + "/x_next"][k]), f"step {k}"
        assert tr["x0"].dtype == torch.float64 and np.array_equal(tr["x0"].numpy(), g[name + "/x0"][k])


@pytest.mark.parametrize("alg,K", [("ddpm", 24), ("ddim", 24), ("ddpm_sympy", 18), ("ddim", 100)])
def test_validate_loops_match_reference(golden_dir, alg, K):
    g = _g(golden_dir, "validate_loop.npz")
    m = _g(golden_dir, "reference_matrices.npz")
    fam = alg.replace("_sympy", "")
    key = f"{alg}_{K:03d}"
    A, B, node = (m[f"{fam}/{key}/{n}"] for n in ("A", "B", "node"))
    net = ToyEps(4, seed=23, out_channels=8)

    def eps_model(z, t):
        ts = torch.ones(z.shape[0], dtype=torch.int32) * int(t)
        return net(z, ts, 0)[:, :4], net(z, ts, 1)[:, :4]

    noise = torch.from_numpy(g[key + "/noise"])
    fresh = [torch.from_numpy(f) for f in g[key + "/fresh"]]
    z, trace = O.validate_ni_loop(A, B, node, eps_model, noise, fresh)
    for k, tr in enumerate(trace):
        assert np.array_equal(tr["x_next"].numpy(), g[key + "/ni_x_next"][k]), f"step {k}"
    if fam == "ddpm":
        zo, _ = O.ddpm_original_loop(K, eps_model, noise, fresh)
    else:
        zo, _ = O.ddim_original_loop(K, eps_model, noise)
    assert np.array_equal(zo.numpy(), g[key + "/original_final"])
    # the reference's executable equivalence: original sampler == NI with the matching matrix
    rel = (z - zo).norm() / zo.norm()
    assert rel < 5e-6, rel


@pytest.mark.parametrize("wname", ["sd3_step_28_weight", "sd3_step_28_weight_sharp"])
@pytest.mark.parametrize("tag,dt", [("f32", torch.float32), ("f16", torch.float16)])
def test_sd3_loop_matches_reference(golden_dir, weights_dir, wname, tag, dt):
    g = _g(golden_dir, "sd3_loop.npz")
    W = O.load_sd3_csv(os.path.join(weights_dir, wname + ".csv"))
    sig = O.sd3_sigmas()
    assert np.array_equal(sig, g["sigmas"])
    net = ToyEps(16, seed=5)
    v_model = lambda x, k: (net(x, 1000 * float(sig[k]), 0), net(x, 1000 * float(sig[k]), 1))
    noises = torch.from_numpy(g[f"{wname}/{tag}/noise"]).to(dt)
    out, trace = O.sd3_ni_loop(W, sig, v_model, noises)
    for k, tr in enumerate(trace):
        assert np.array_equal(tr["x_in"].float().numpy(), g[f"{wname}/{tag}/x_in"][k]), f"x_in step {k}"
        assert np.array_equal(tr["out"].float().numpy(), g[f"{wname}/{tag}/out"][k]), f"out step {k}"


def test_sd3_functions_match_reference(golden_dir):
    g = _g(golden_dir, "sd3_loop.npz")
    xs = [torch.from_numpy(x) for x in g["fn/xs"]]
    assert np.array_equal(O.sd3_weighted_sum(xs, None).numpy(), g["fn/uniform_mean"])
    seq = [[float(w), x] for w, x in zip(g["fn/euler_w"], xs)]
    acc, eq = O.euler_weighted_sum(seq, 0)
    assert np.array_equal(acc.numpy(), g["fn/euler_acc"]) and np.array_equal(eq.numpy(), g["fn/euler_equiv"])
    assert np.array_equal(O.euler_weighted_sum(seq, 3)[1].numpy(), g["fn/euler_clip3_equiv"])


def test_common_form_equals_sd3_loop(weights_dir):
    """(A,B) conversion of the csv table reproduces the SD3 loop: the model input of
    step k+1 is row k of [A|B], and the final output is row K-1 (sigma_K = 0)."""
    W = O.load_sd3_csv(os.path.join(weights_dir, "sd3_step_28_weight_sharp.csv"))
    sig = O.sd3_sigmas()
    A, B, node = O.sd3_triple(W, sig)
    net = ToyEps(16, seed=5)
    v_model = lambda x, k: (net(x, 1000 * float(sig[k]), 0), net(x, 1000 * float(sig[k]), 1))
    g = torch.Generator().manual_seed(10)
    noises = torch.randn(2, 16, 8, 8, generator=g)
    out, trace = O.sd3_ni_loop(W, sig, v_model, noises)
    x = noises.double()
    hist = []
    for k in range(28):
        v_text, v_null = v_model(x.float(), k)
        s = float(sig[k])
        x0, x = O.ni_step_f64(1.0, (-s * 7, -s * (1 - 7)), x, (v_text, v_null), A[k], hist, B[k], [noises])
        hist.append(x0)
        if k < 27:
            ref = trace[k + 1]["x_in"]
        else:
            ref = out
        assert (x.float() - ref).abs().max() / ref.norm() < 1e-5
