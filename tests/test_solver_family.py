"""CPU: the original samplers that head the reference's FID tables (results/FID/{dpmsolver,dpmsolverpp,deis}_*step.csv) as
coefficient matrices.  The generators (product, coefficient space) and the oracle's tensor-level
restatements of the ORIGINAL loops are written independently; both are pinned here:
  * DPM-Solver / DPM-Solver++ multistep-2/3 and singlestep-2/3 on the 5/10/15-step quadratic grids against matrices that
    tests/golden/make_golden.py obtained from the reference's unmodified DPM_Solver class (solver_matrices.npz);
  * DEIS t-AB / rho-AB / rho-RK / iPNDM against matrices obtained from the reference's unmodified th_deis executed with a
    numpy-backed stand-in for jax (deis_matrices.npz), against each other and against the classical Adams-Bashforth /
    Runge-Kutta tables they reduce to."""
import os

import numpy as np
import pytest
import torch

from naturaldiffusion_b200 import generators as G
from oracle import ni_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solver_matrices.npz")


def _settings(z):
    """(key, K, algorithm, method, order, skip_type, lower_order_final) of every matrix in the golden file: the 24 FID-table
    settings (5/10/15 steps, time_quadratic) plus the other time grids, the fixed-order singlestep driver, the
    lower_order_final tail and first order"""
    out = []
    for key in sorted(set(k.rsplit("/", 1)[0] for k in z.files)):
        parts = key.split("/")
        skip, lof = ("time_quadratic", False) if len(parts) == 3 else (parts[3], parts[4] == "lof1")
        out.append((key, int(parts[2]), parts[0], parts[1][:-1], int(parts[1][-1]), skip, lof))
    return out


def test_dpm_solver_generators_match_the_reference_solver_class():
    z = np.load(GOLD)
    settings = _settings(z)
    assert len(settings) == 39
    for key, K, alg, method, order, skip, lof in settings:
        t = G.dpm_solver_triple(K, alg, method, order, skip_type=skip, lower_order_final=lof)
        scale = max(1.0, np.abs(z[key + "/A"]).max())
        assert np.abs(t.A - z[key + "/A"]).max() < 1e-11 * scale, key
        assert np.abs(t.B - z[key + "/B"]).max() < 1e-11 * scale, key
        assert np.abs(t.node - z[key + "/node"]).max() < 1e-12, key
        assert np.all(t.B[:, 1:] == 0)  # deterministic samplers: only the initial noise


def test_oracle_dpm_solver_restatement_matches_the_reference_solver_class():
    """the oracle's tensor-level sample() run in coefficient space (fp64) reproduces the same goldens"""
    z = np.load(GOLD)
    for key, K, alg, method, order, skip, lof in _settings(z):
        if method == "singlestep_fixed":
            continue  # the oracle restates the two drivers the FID runs use
        ns = O._VPSchedule(torch.float64)
        rows, calls = [], [0]

        def eps_model(x, t):
            if calls[0] > 0:
                rows.append(x[0].clone())
            y = torch.zeros_like(x)
            y[0, calls[0]] = 1.0
            calls[0] += 1
            return (x - ns.alpha(t) * y) / ns.std(t)

        x0 = torch.zeros(1, 2 * K + 1, dtype=torch.float64)
        x0[0, K] = 1.0
        xe = O.dpm_solver_original_sample(eps_model, x0, K, alg, method, order, skip_type=skip, lower_order_final=lof, dtype=torch.float64)
        rows.append(xe[0])
        M = torch.stack(rows).numpy()
        scale = max(1.0, np.abs(z[key + "/A"]).max())
        assert calls[0] == K and np.abs(M[:, :K] - z[key + "/A"]).max() < 1e-11 * scale and np.abs(M[:, K:] - z[key + "/B"]).max() < 1e-11 * scale, key


def _apply_matrix(t, eps_model, noise):
    """x_{k+1} = sum_j A[k,j] x0_j + B[k,0] eps_0 with x0_k = (x_k - sigma_k eps)/alpha_k, plain torch fp64"""
    ts, al, sg = t.node[:, 0], t.node[:, 1], t.node[:, 2]
    x, x0s = noise.clone(), []
    for k in range(t.K):
        x0s.append((x - sg[k] * eps_model(x, float(ts[k]))) / al[k])
        x = sum(t.A[k, j] * x0s[j] for j in range(k + 1)) + t.B[k, 0] * noise
    return x


@pytest.mark.parametrize("method,kw", [("rho_ab", dict(ab_order=3)), ("rho_ab", dict(ab_order=2)), ("rho_rk", dict()),
                                       ("rho_rk", dict(rk_method="2heun")), ("ipndm", dict()), ("rho_ab", dict(ts_phase="rho")),
                                       ("rho_rk", dict(ts_phase="rho"))])
def test_deis_matrices_equal_the_original_loops(method, kw):
    torch.manual_seed(0)
    noise = torch.randn(4, 3, 8, 8, dtype=torch.float64)
    sgm = lambda t: float(np.sqrt(1 - O._deis_abar(t)))
    eps_model = lambda x, t: sgm(t) * x + 0.1 * torch.tanh(1.3 * x + t)
    for n in (5, 10, 15):
        t = G.deis_triple(n, method, **kw)
        a = _apply_matrix(t, eps_model, noise)
        b = O.deis_original_sample(eps_model, noise, n, method, **kw)
        assert float((a - b).abs().max() / b.norm()) < 2e-6, (method, kw, n)


DEIS_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deis_matrices.npz")


def _deis_settings(z):
    out = []
    for key in sorted(set(k.rsplit("/", 1)[0] for k in z.files)):
        m, n, o, ph, rk = key.split("/")
        out.append((key, m, int(n), int(o[5:]), ph, rk))
    return out


def test_deis_generators_match_the_reference_th_deis():
    """38 settings traced from the reference's own th_deis.get_sampler (executed with the numpy stand-in for jax): t_ab / rho_ab
    order 2 and 3 and rho_rk (Kutta) at 5/10/15 steps on the `t` and `rho` grids -- the rows of results/FID/deis_*step.csv --
    plus iPNDM, other tableaus and the `log` grid"""
    z = np.load(DEIS_GOLD)
    settings = _deis_settings(z)
    assert len(settings) == 38
    for key, m, n, o, ph, rk in settings:
        t = G.deis_triple(n, m, ab_order=o, rk_method=rk, ts_phase=ph)
        scale = max(1.0, np.abs(z[key + "/A"]).max())
        assert np.abs(t.A - z[key + "/A"]).max() < 1e-11 * scale, key
        assert np.abs(t.B - z[key + "/B"]).max() < 1e-11 * scale, key
        assert np.abs(t.node[:-1] - z[key + "/node"]).max() < 1e-12, key


def test_oracle_deis_loops_match_the_reference_th_deis():
    """the oracle's tensor-level DEIS loops run in coefficient space (fp64) against the same goldens"""
    z = np.load(DEIS_GOLD)
    for key, m, n, o, ph, rk in _deis_settings(z):
        if rk not in ("3kutta", "2heun", "4rk", "3heun"):
            continue
        K = z[key + "/A"].shape[0]
        rows, calls = [], [0]

        def eps_model(x, t):
            if calls[0] > 0:
                rows.append(x[0].clone())
            ab = float(O._deis_abar(t))
            y = torch.zeros_like(x)
            y[0, calls[0]] = 1.0
            calls[0] += 1
            return (x - np.sqrt(ab) * y) / np.sqrt(1.0 - ab)

        x0 = torch.zeros(1, 2 * K + 1, dtype=torch.float64)
        x0[0, K] = 1.0
        if m == "t_ab":
            xe = O.deis_tab_original_loop(O.deis_rev_ts(n, 2, ph), eps_model, x0, ab_order=o)
        else:
            xe = O.deis_original_sample(eps_model, x0, n, m, ab_order=o, rk_method=rk, ts_phase=ph)
        rows.append(xe[0])
        M = torch.stack(rows).numpy()
        scale = max(1.0, np.abs(z[key + "/A"]).max())
        assert calls[0] == K, key
        assert np.abs(M[:, :K] - z[key + "/A"]).max() < 1e-10 * scale and np.abs(M[:, K:] - z[key + "/B"]).max() < 1e-10 * scale, key


def test_deis_building_blocks_reduce_to_the_classical_tables():
    # Adams-Bashforth weights on a uniform grid: h * (23, -16, 5)/12 and h * (3, -1)/2 (Riemann-sum accuracy 1/10000)
    grid = np.linspace(3.0, 1.0, 9)
    h = grid[1] - grid[0]
    C = G._ab_coefficients(grid, lambda a, b: np.ones_like(a), lambda a: np.ones_like(a), 3)
    assert np.allclose(C[0, :1], [h], rtol=1e-3) and np.allclose(C[1, :2], [1.5 * h, -0.5 * h], rtol=1e-3)
    assert np.allclose(C[2, :3], np.array([23, -16, 5]) * h / 12, rtol=1e-3)
    assert np.allclose(C[5], np.array([55, -59, 37, -9]) * h / 24, rtol=1e-3)
    assert np.allclose(O.deis_rho_ab_coefficients(grid, 3), C, rtol=1e-12, atol=1e-15)
    # Runge-Kutta tableaus: consistency conditions
    for name, (c, a, b) in G.RK_TABLEAUS.items():
        assert abs(sum(b) - 1) < 1e-12, name
        for ci, row in zip(c, a):
            assert abs(ci - sum(row)) < 1e-12, name
    # the t_ab special case is the generator that reproduces the shipped results/deis matrices
    a, b = G.deis_triple(10, "t_ab"), G._deis_tab_on(G._DeisVP().rev_ts(10), 3, 10000)
    assert np.abs(a.A - b.A).max() < 1e-14
    # time grids agree between product and oracle
    for phase in ("t", "rho", "log"):
        assert np.allclose(G._DeisVP().rev_ts(10, 2, phase), O.deis_rev_ts(10, 2, phase), rtol=1e-12)


def test_third_order_solvers_converge_faster_than_first_order():
    """sanity of the whole family on a Gaussian-data problem with a closed-form eps: x0 ~ N(0, s^2) => eps(x,t) linear in x"""
    s2 = 0.25
    vp = G.VPLinearSchedule()

    def final_std(t):  # std of x_K when x_T ~ N(0,1): |sum of the linear map| -- exact answer is sqrt(alpha0^2 s2 + sigma0^2)
        al, sg = t.node[:, 1], t.node[:, 2]
        x, x0s = 1.0, []
        for k in range(t.K):
            eps = sg[k] * x / (al[k] ** 2 * s2 + sg[k] ** 2)
            x0s.append((x - sg[k] * eps) / al[k])
            x = sum(t.A[k, j] * x0s[j] for j in range(k + 1)) + t.B[k, 0] * 1.0
        return abs(x)

    # the initial state of the true process has variance alpha_T^2 s2 + sigma_T^2 (not 1): compare against the exact flow of x_T = 1
    def exact(t_end=1e-3):
        v = lambda t: vp.alpha(t) ** 2 * s2 + vp.sigma(t) ** 2
        return np.sqrt(v(t_end) / v(1.0))

    errs = {name: abs(final_std(t) - exact()) for name, t in
            dict(o1=G.dpm_solver_triple(15, "dpmsolver++", "multistep", 1), o2=G.dpm_solver_triple(15, "dpmsolver++", "multistep", 2),
                 o3=G.dpm_solver_triple(15, "dpmsolver++", "multistep", 3), rk=G.deis_triple(5, "rho_rk"), ab=G.deis_triple(15, "rho_ab")).items()}
    assert errs["o3"] < errs["o2"] < errs["o1"] and errs["rk"] < errs["o1"] and errs["ab"] < errs["o1"], errs
