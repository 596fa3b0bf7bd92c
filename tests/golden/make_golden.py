"""Generate tests/golden/*.npz by running the REFERENCE'S OWN functions (imported from
/root/reference with third-party modules stubbed -- oracle/ref_loader.py) on small seeded inputs.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The outputs are committed; tests/test_oracle_golden.py checks the oracle restatement against them
and tests/test_gpu_parity.py checks the CUDA path against them.
"""
import glob
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

os.environ["NI_EXEC_REFERENCE"] = "1"  # generating fixtures is the one place that executes reference code
from oracle import ref_loader  # noqa: E402
from toy_models import ToyEps  # noqa: E402

REF = ref_loader.REF_ROOT
assert ref_loader.available(), "reference tree not found"
torch.set_grad_enabled(False)


def gen_cifar():
    """src/CIFAR10NaturalInference.py:292-306 with the reference's data_fn / weighted_sum / VP score wrapper."""
    data_fn, weighted_sum = ref_loader.cifar_functions()
    VPSDE, get_score_fn = ref_loader.score_sde_vp()
    sde = VPSDE(beta_min=0.1, beta_max=20, N=1000)
    out = {}
    for name in ("step_5_weight_00", "step_10_weight_42", "step_15_weight_173"):
        A, B, node = np.load(os.path.join(REF, "weights", name + ".npz")).values()
        net = ToyEps(3, seed=11)

        class M(torch.nn.Module):
            def forward(self, x, labels):
                return net(x, labels)

        score_fn = get_score_fn(sde, M(), train=False, continuous=True)
        torch.manual_seed(888)
        noise = torch.randn(6, 3, 8, 8, dtype=torch.float32)
        ts = node[:, 0]
        seq_x0, x = [], noise
        xs, x0s = [], []
        for kk in range(ts.shape[0] - 1):
            pred_x0 = data_fn(score_fn, x, ts[kk], node[kk, 1], node[kk, 2], "cpu")
            seq_x0.append(pred_x0)
            next_x0 = weighted_sum(A[kk], seq_x0)
            next_eps = B[kk, 0] * noise
            x = next_x0 + next_eps
            xs.append(x.numpy().copy())
            x0s.append(pred_x0.numpy().copy())
        out[name + "/noise"] = noise.numpy()
        out[name + "/x_next"] = np.stack(xs)
        out[name + "/x0"] = np.stack(x0s)  # float64, as the reference keeps it
    np.savez_compressed(os.path.join(HERE, "cifar_loop.npz"), **out)


def gen_validate():
    """src/ValidateNaturalInference.py: original DDPM/DDIM loops (:235-250, :288-302) and the NI loop
    (:349-366) with the reference's own coefficient tables, calc_x0_mean_z and weighted_sum; toy CFG model."""
    v = ref_loader.validate_module()
    out = {}
    net = ToyEps(4, seed=23, out_channels=8)

    def forward_cfg(z, timesteps):
        cond = net(z, timesteps, 0)[:, :4]
        uncond = net(z, timesteps, 1)[:, :4]
        return cond, uncond, uncond + 4.0 * (cond - uncond)

    for alg, K in (("ddpm", 24), ("ddim", 24), ("ddpm_sympy", 18), ("ddim", 100)):
        fam = alg.replace("_sympy", "")
        A, B, node = np.load(os.path.join(REF, "results", fam, "%s_%03d.npz" % (alg, K))).values()
        torch.manual_seed(0)
        n = 3
        noise = torch.randn(n, 4, 8, 8)
        fresh = [torch.randn(n, 4, 8, 8) for _ in range(K)]
        # --- natural inference (:321-366)
        coeff_all, skip_idxs = v.skip_ddim_coeff(v.create_ddim_coeff(), K)
        coeff_all = [torch.from_numpy(e).to(dtype=torch.float32) for e in coeff_all]
        c1, c2 = coeff_all[2].flip(0), coeff_all[3].flip(0)
        seq_x0, seq_eps = [], [noise]
        z = noise.clone()
        zs = []
        for kk in range(K):
            ts = torch.ones(n, dtype=torch.int32) * int(node[kk, 0])
            cond, uncond, fuse = forward_cfg(z, ts)
            pred_x0 = c1[kk] * z - c2[kk] * fuse
            seq_x0.append(pred_x0)
            seq_eps.append(fresh[kk])
            z = v.weighted_sum(A[kk], seq_x0) + v.weighted_sum(B[kk], seq_eps)
            zs.append(z.numpy().copy())
        key = "%s_%03d" % (alg, K)
        out[key + "/noise"] = noise.numpy()
        out[key + "/fresh"] = np.stack([f.numpy() for f in fresh])
        out[key + "/ni_x_next"] = np.stack(zs)
        # --- original sampler (:213-250 / :268-302)
        if fam == "ddpm":
            ca, idxs = v.skip_ddpm_coeff(v.create_ddpm_coeff(), K)
            ca = [torch.from_numpy(e).to(dtype=torch.float32) for e in ca]
            alphas, abar, log_var, cxt2x0, ceps2x0, cxt, cx0 = ca
        else:
            ca, idxs = v.skip_ddim_coeff(v.create_ddim_coeff(), K)
            ca = [torch.from_numpy(e).to(dtype=torch.float32) for e in ca]
            alphas, abar, cxt2x0, ceps2x0, cxt, cx0 = ca
        coeff = cxt2x0, ceps2x0, cxt, cx0
        z = noise.clone()
        for m, ii in enumerate(list(range(K))[::-1]):
            ts = torch.ones(n, dtype=torch.int32) * idxs[ii]
            cond, uncond, fuse = forward_cfg(z, ts)
            x0, mean_z = v.calc_x0_mean_z(z, fuse, coeff, ii)
            z = mean_z + torch.exp(0.5 * log_var[ii]) * fresh[m] if fam == "ddpm" else mean_z
        out[key + "/original_final"] = z.numpy()
    np.savez_compressed(os.path.join(HERE, "validate_loop.npz"), **out)


def gen_sd3():
    """src/SD3NaturalInference.py: weighted_sum (:157-168), euler_weighted_sum (:61-69) and the loop (:198-223)
    in fp32 and in the reference's fp16, toy velocity model; sigmas from the restated scheduler formula."""
    s = ref_loader.sd3_module()
    import pandas as pd
    from oracle.ni_oracle import sd3_sigmas
    sig = torch.from_numpy(sd3_sigmas())
    out = {"sigmas": sig.numpy()}
    net = ToyEps(16, seed=5)
    for wname in ("sd3_step_28_weight", "sd3_step_28_weight_sharp"):
        W = pd.read_csv(os.path.join(REF, "weights", wname + ".csv"), index_col=0).to_numpy()
        for dt, tag in ((torch.float32, "f32"), (torch.float16, "f16")):
            g = torch.Generator().manual_seed(10)
            noises = torch.randn(2, 16, 8, 8, generator=g).to(dt)
            seq, xins, outs = [], [], []
            for kk in range(28):
                sigma = sig[kk]
                curr = s.weighted_sum(seq, W) if len(seq) != 0 else torch.zeros_like(noises)
                x_in = sigma * noises + (1 - sigma) * curr
                v_text = net(x_in, 1000 * float(sigma), 0)
                v_null = net(x_in, 1000 * float(sigma), 1)
                x0_null = x_in - sigma * v_null
                x0_text = x_in - sigma * v_text
                x0 = x0_null + 7 * (x0_text - x0_null)
                seq.append(x0)
                o = s.weighted_sum(seq, W)
                xins.append(x_in.float().numpy().copy())
                outs.append(o.float().numpy().copy())
            out[f"{wname}/{tag}/noise"] = noises.float().numpy()
            out[f"{wname}/{tag}/x_in"] = np.stack(xins)
            out[f"{wname}/{tag}/out"] = np.stack(outs)
    # stand-alone function vectors
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(2, 16, 4, 4, generator=g) for _ in range(5)]
    out["fn/xs"] = np.stack([x.numpy() for x in xs])
    out["fn/uniform_mean"] = s.weighted_sum(xs, None).numpy()
    ws = [0.3, 0.1, 0.25, 0.05, 0.3]
    acc, eq = s.euler_weighted_sum([[w, x] for w, x in zip(ws, xs)], 0)
    out["fn/euler_w"] = np.array(ws)
    out["fn/euler_acc"], out["fn/euler_equiv"] = acc.numpy(), eq.numpy()
    acc3, eq3 = s.euler_weighted_sum([[w, x] for w, x in zip(ws, xs)], 3)
    out["fn/euler_clip3_equiv"] = eq3.numpy()
    np.savez_compressed(os.path.join(HERE, "sd3_loop.npz"), **out)


def gen_weights():
    """The shipped inputs of the hot path (weights/*.npz, weights/*.csv): re-saved array by array / line by line so the
    fixtures carry exactly the reference's numbers (positional npz order kept: xstart, epsilon, node)."""
    out_dir = os.path.join(os.path.dirname(os.path.dirname(HERE)), "naturaldiffusion_b200", "data", "weights")
    os.makedirs(out_dir, exist_ok=True)
    for f in sorted(glob.glob(os.path.join(REF, "weights", "*.npz"))):
        with np.load(f) as z:
            np.savez(os.path.join(out_dir, os.path.basename(f)), **{k: z[k] for k in z.files})
    for f in sorted(glob.glob(os.path.join(REF, "weights", "*.csv"))):
        with open(f) as src, open(os.path.join(out_dir, os.path.basename(f)), "w") as dst:
            dst.write(src.read())


def gen_config_matrices():
    """The two matrices BASELINE's configs need but the reference does not ship: ddim_010 (C1) and ddpm_250 (C4), produced by
    the reference's OWN generators (src/AnalyzeDDPMDDIM.py `ddim_analyze_coeff`, `ddpm_analyze_coeff`) with their file
    writer intercepted (the reference tree is read-only)."""
    import contextlib
    import io
    mod = ref_loader.analyze_ddpm_ddim_module()
    got = {}

    def capture(A, B, node, out_dir, prefix):
        got["%s_%03d" % (prefix, A.shape[0])] = (A.copy(), B.copy(), node.copy())

    mod.save_coeff_matrix = capture
    with contextlib.redirect_stdout(io.StringIO()):
        mod.ddim_analyze_coeff(10)
        mod.ddpm_analyze_coeff(250)
        mod.ddpm_analyze_coeff(10)
    out = {}
    for key, (A, B, node) in got.items():
        out[key + "/A"], out[key + "/B"], out[key + "/node"] = A, B, node
    np.savez_compressed(os.path.join(HERE, "config_matrices.npz"), **out)


def gen_matrices():
    """The shipped coefficient matrices (results/*/*.npz): golden outputs of the reference's generators.
    K <= 201 are stored whole; the two K=500 files as float64 row/column digests to keep the repo small."""
    out = {}
    for f in sorted(glob.glob(os.path.join(REF, "results", "*", "*.npz"))):
        A, B, node = np.load(f).values()
        key = os.path.relpath(f, os.path.join(REF, "results"))[:-4]
        if A.shape[0] <= 201:
            out[key + "/A"], out[key + "/B"], out[key + "/node"] = A, B, node
        else:
            out[key + "/digest"] = np.concatenate([A.sum(0), A.sum(1), B.sum(0), B.sum(1), np.diag(A), node.ravel()])
    np.savez_compressed(os.path.join(HERE, "reference_matrices.npz"), **out)


def gen_solver_matrices():
    """Coefficient matrices of the ORIGINAL DPM-Solver / DPM-Solver++ samplers, obtained by running the reference's own
    `DPM_Solver.sample` (deps/dpm_solver_pytorch.py, unmodified) in coefficient space: x lives in R^(2K+1) over the basis
    (y_0..y_{K-1}, eps_0..eps_K), the j-th model call returns (x - alpha y_j)/sigma and records x as row j-1 of [A|B]
    (what src/AnalyzeDPMSolver.py does with sympy).  Same arguments as the FID runs of
    src/CIFAR10NaturalInference.py:363-393 (time_quadratic, lower_order_final=False, denoise_to_zero=False), in float64
    (the reference's float32 time grid is a precision detail, not part of the sampler)."""
    dpm = ref_loader.dpm_solver_module()
    torch.set_default_dtype(torch.float64)
    out = {}
    try:
        settings = [(K, alg, method, order, "time_quadratic", False) for K in (5, 10, 15) for alg in ("dpmsolver", "dpmsolver++")
                    for method, order in (("multistep", 2), ("multistep", 3), ("singlestep", 2), ("singlestep", 3))]
        # beyond the FID-table settings: the other time grids of get_time_steps, the fixed-order singlestep driver, the
        # lower_order_final tail (the library default) and first order (= DDIM)
        settings += [(10, alg, method, 3, skip, False) for alg in ("dpmsolver", "dpmsolver++") for method in ("multistep", "singlestep")
                     for skip in ("time_uniform", "logSNR")]
        settings += [(9, "dpmsolver++", "singlestep_fixed", 3, "time_uniform", False), (8, "dpmsolver", "singlestep_fixed", 2, "logSNR", False),
                     (7, "dpmsolver++", "multistep", 3, "time_uniform", True), (6, "dpmsolver", "multistep", 2, "time_quadratic", True),
                     (10, "dpmsolver++", "multistep", 1, "time_uniform", False), (11, "dpmsolver", "singlestep", 3, "time_quadratic", False),
                     (13, "dpmsolver++", "singlestep", 3, "time_quadratic", False)]
        for K, alg, method, order, skip, lof in settings:
            if True:
                if True:
                    ns = dpm.NoiseScheduleVP("linear", continuous_beta_0=0.1, continuous_beta_1=20.0)
                    rows, nodes = [], []

                    def noise_fn(x, t):
                        tt = t.reshape(-1)[:1]
                        if nodes:
                            rows.append(x[0].clone())
                        j = len(nodes)
                        nodes.append([float(tt), float(ns.marginal_alpha(tt)), float(ns.marginal_std(tt))])
                        y = torch.zeros_like(x)
                        y[0, j] = 1.0
                        return (x - ns.marginal_alpha(tt) * y) / ns.marginal_std(tt)

                    solver = dpm.DPM_Solver(noise_fn, ns, algorithm_type=alg)
                    x0 = torch.zeros(1, 2 * K + 1)
                    x0[0, K] = 1.0
                    xe = solver.sample(x0, steps=K, t_start=1.0, t_end=1e-3, order=order, skip_type=skip, method=method,
                                       denoise_to_zero=False, lower_order_final=lof)
                    rows.append(xe[0].clone())
                    t_end = torch.tensor([1e-3])
                    nodes.append([1e-3, float(ns.marginal_alpha(t_end)), float(ns.marginal_std(t_end))])
                    assert len(rows) == K and len(nodes) == K + 1, (K, alg, method, order, len(rows), len(nodes))
                    M = torch.stack(rows).numpy()
                    key = f"{alg}/{method}{order}/{K:03d}" if (skip == "time_quadratic" and not lof) else f"{alg}/{method}{order}/{K:03d}/{skip}/lof{int(lof)}"
                    out[key + "/A"], out[key + "/B"], out[key + "/node"] = M[:, :K], M[:, K:], np.array(nodes)
    finally:
        torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, "solver_matrices.npz"), **out)


def gen_deis_matrices():
    """Coefficient matrices of the ORIGINAL DEIS samplers, obtained by running the reference's own `th_deis.get_sampler`
    (deps/th_deis, unmodified) in coefficient space, with oracle/jax_numpy_shim.py standing in for jax (not installed here):
    numpy float64 arithmetic, complex-step grad, loop vmap.  Settings of src/CIFAR10NaturalInference.py:166-178 (the rows of
    results/FID/deis_*step.csv: ts_phase t | rho, ts_order 2, t_ab | rho_ab | rho_rk, ab_order 2 | 3) plus iPNDM, Heun's
    second-order RK and the `log` time grid."""
    tdeis = ref_loader.th_deis_module()
    t2a, a2t = tdeis.get_linear_alpha_fns(0.1, 20.0)
    sde = tdeis.VPSDE(t2a, a2t, 1e-3, 1.0)
    stages = {"3kutta": 3, "2heun": 2, "4rk": 4, "3heun": 3}
    settings = [(n, m, o, ph, "3kutta") for n in (5, 10, 15) for ph in ("t", "rho") for m in ("t_ab", "rho_ab") for o in (2, 3)]
    settings += [(n, "rho_rk", 3, ph, "3kutta") for n in (5, 10, 15) for ph in ("t", "rho")]
    settings += [(n, "ipndm", 3, "t", "3kutta") for n in (5, 10, 15)]
    settings += [(6, "rho_rk", 3, "t", "2heun"), (4, "rho_rk", 3, "t", "4rk"), (5, "rho_rk", 3, "log", "3heun"), (8, "rho_ab", 3, "log", "3kutta"),
                 (8, "t_ab", 1, "t", "3kutta")]
    out = {}
    for num_step, method, ab_order, ts_phase, rk in settings:
        nfe = num_step * stages[rk] if method == "rho_rk" else num_step
        rows, nodes = [], []

        def eps_fn(x, t):
            tt = float(t)
            if nodes:
                rows.append(x[0].clone())
            j = len(nodes)
            ab = float(t2a(np.float64(tt)))
            nodes.append([tt, np.sqrt(ab), np.sqrt(1.0 - ab)])
            y = torch.zeros_like(x)
            y[0, j] = 1.0
            return (x - np.sqrt(ab) * y) / np.sqrt(1.0 - ab)

        x0 = torch.zeros(1, 2 * nfe + 1, dtype=torch.float64)
        x0[0, nfe] = 1.0
        xe = tdeis.get_sampler(sde, eps_fn, ts_phase, 2, num_step, method=method, ab_order=ab_order, rk_method=rk)(x0)
        rows.append(xe[0].clone())
        assert len(rows) == nfe and len(nodes) == nfe, (method, num_step, len(rows), len(nodes))
        M = torch.stack(rows).numpy()
        key = f"{method}/{num_step:03d}/order{ab_order}/{ts_phase}/{rk}"
        out[key + "/A"], out[key + "/B"], out[key + "/node"] = M[:, :nfe], M[:, nfe:], np.array(nodes)
    np.savez_compressed(os.path.join(HERE, "deis_matrices.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "solvers":
        gen_solver_matrices()
        gen_deis_matrices()
        raise SystemExit(0)
    gen_solver_matrices()
    gen_deis_matrices()
    gen_weights()
    gen_cifar()
    gen_validate()
    gen_sd3()
    gen_matrices()
    gen_config_matrices()
    for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        print(os.path.basename(f), os.path.getsize(f))
