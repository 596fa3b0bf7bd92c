"""Coefficient-matrix generators for first-order samplers, host-side numpy.

The reference ships DDPM/DDIM matrices only for K in {18, 24, 100, 500} and flow-Euler for {18, 24};
BASELINE's configs need ddim_010 (C1) and ddpm_250 (C4).  These are produced here from the same closed
forms as src/AnalyzeDDPMDDIM.py:126-174 (`ddpm_analyze_coeff`), :297-340 (`ddim_analyze_coeff`) and
src/AnalyzeFlowMatching.py:20-59 (`flow_analyze_coeff`), but built row by row from the Markov property of
first-order samplers  row_k[:k] = p_k * row_{k-1}[:k]  instead of re-multiplying every product.
Checked against the shipped matrices in tests/test_host_logic.py (<= 1e-14).
"""
from __future__ import annotations

import numpy as np

from .coeffs import CoeffTriple, spaced_timesteps


def first_order_triple(p, q, r, node, name="") -> CoeffTriple:
    """x_{k+1} = p_k x_k + q_k x0_k + r_k eps_{k+1} (sampling order k = 0..K-1, x_0 = eps_0)  ->  (A, B, node)."""
    p, q = np.asarray(p, np.float64), np.asarray(q, np.float64)
    K = len(p)
    r = np.zeros(K) if r is None else np.asarray(r, np.float64)
    A = np.zeros((K, K))
    B = np.zeros((K, K + 1))
    prev_a, prev_b = np.zeros(K), np.zeros(K + 1)
    prev_b[0] = 1.0
    for k in range(K):
        a, b = p[k] * prev_a, p[k] * prev_b
        a[k] = q[k]
        b[k + 1] = r[k]
        A[k], B[k] = a, b
        prev_a, prev_b = a, b
    return CoeffTriple(A, B, node, name=name)


def _vp_discrete(num_step: int):
    idx = np.array(spaced_timesteps(1000, num_step))
    ab = np.cumprod(1.0 - np.linspace(0.0001, 0.02, 1000, dtype=np.float64))[idx]
    ab_prev = np.append(1.0, ab[:-1])
    node = np.zeros((num_step + 1, 3))
    node[0] = [999, 0.0, 1.0]                      # the reference's convention for the start node
    node[1:num_step, 0] = idx[:-1][::-1][: num_step - 1]
    node[1:num_step, 1] = np.sqrt(ab[:-1])[::-1]
    node[1:num_step, 2] = np.sqrt(1 - ab[:-1])[::-1]
    node[num_step] = [-1, 1.0, 0.0]
    return idx, ab, ab_prev, node


def ddim_triple(num_step: int) -> CoeffTriple:
    idx, ab, ab_prev, node = _vp_discrete(num_step)
    rect = np.sqrt((1 - ab_prev) / (1 - ab))
    return first_order_triple(rect[::-1], (np.sqrt(ab_prev) - rect * np.sqrt(ab))[::-1], None, node, name=f"ddim_{num_step:03d}")


def ddpm_triple(num_step: int) -> CoeffTriple:
    idx, ab, ab_prev, node = _vp_discrete(num_step)
    alphas = ab / ab_prev
    betas = 1 - alphas
    var = betas * (1 - ab_prev) / (1 - ab)
    std = np.sqrt(np.exp(np.log(np.append(1e-5, var[1:]))))
    cx0 = np.sqrt(ab_prev) * betas / (1 - ab)
    cxt = np.sqrt(alphas) * (1 - ab_prev) / (1 - ab)
    return first_order_triple(cxt[::-1], cx0[::-1], std[::-1], node, name=f"ddpm_{num_step:03d}")


def flow_euler_triple(num_step: int, sigmas=None) -> CoeffTriple:
    """Flow-matching Euler on sigma grid (default linspace(1, 0, K+1), the reference's); x = (1-s) x0 + s eps."""
    sig = np.linspace(1, 0, num_step + 1) if sigmas is None else np.asarray(sigmas, np.float64)
    p = sig[1:] / sig[:-1]
    node = np.stack([sig, 1 - sig, sig], axis=1)
    return first_order_triple(p, 1 - p, None, node, name=f"flow_euler_{num_step:03d}")


def markov_ratio(triple: CoeffTriple, tol: float = 1e-12):
    """see coeffs.markov_ratios (kept here for discoverability next to the generators)"""
    from .coeffs import markov_ratios
    mr = markov_ratios(triple, tol)
    return None if mr is None else np.array(mr[0])


# ----------------------------------------------------------------------------------------------
# any linear sampler -> matrices, by running it in coefficient space
# ----------------------------------------------------------------------------------------------

class VPLinearSchedule:
    """Continuous VP schedule beta(t) = beta_0 + t (beta_1 - beta_0): log alpha, alpha, sigma, lambda = log(alpha/sigma)
    and its inverse (the `NoiseScheduleVP('linear')` of deps/dpm_solver_pytorch.py:6-167, float64 numpy)."""

    def __init__(self, beta_0=0.1, beta_1=20.0):
        self.b0, self.b1 = beta_0, beta_1

    def log_alpha(self, t):
        return -0.25 * t ** 2 * (self.b1 - self.b0) - 0.5 * t * self.b0

    def alpha(self, t):
        return np.exp(self.log_alpha(t))

    def sigma(self, t):
        return np.sqrt(1.0 - np.exp(2.0 * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * np.log(1.0 - np.exp(2.0 * la))

    def inv_lam(self, lam):
        tmp = 2.0 * (self.b1 - self.b0) * np.logaddexp(-2.0 * lam, 0.0)
        return tmp / (np.sqrt(self.b0 ** 2 + tmp) + self.b0) / (self.b1 - self.b0)


class CoefficientTracer:
    """Run an (unmodified, linear) sampler on vectors of R^{2K+1} over the basis (y_0..y_{K-1}, eps_0..eps_K):
    `model(x, t)` hands back the unit vector of the next x0-prediction and records x as a matrix row -- the state a
    sampler feeds to its j-th model call IS row j-1 of [A|B] -- `noise()` hands back the next noise unit vector.
    What the reference does with sympy symbols (src/AnalyzeDPMSolver.py:272-281, src/AnalyzeDEIS.py:80-87), done with
    plain linear algebra."""

    def __init__(self, num_calls: int, schedule=None):
        self.K, self.ns = num_calls, schedule
        self.rows, self.nodes, self.calls, self.draws = [], [], 0, 0

    def unit(self, i):
        v = np.zeros(2 * self.K + 1)
        v[i] = 1.0
        return v

    def noise(self):
        v = self.unit(self.K + self.draws)
        self.draws += 1
        return v

    def _node(self, t):
        self.nodes.append([t, self.ns.alpha(t), self.ns.sigma(t)] if self.ns is not None else [t, np.nan, np.nan])

    def model_x0(self, x, t):
        """data-prediction call"""
        if self.calls > 0:
            self.rows.append(np.array(x, dtype=np.float64))
        self._node(t)
        self.calls += 1
        return self.unit(self.calls - 1)

    def model_eps(self, x, t):
        """noise-prediction call: eps = (x - alpha y)/sigma with y the new x0 symbol"""
        y = self.model_x0(x, t)
        return (np.asarray(x) - self.ns.alpha(t) * y) / self.ns.sigma(t)

    def finish(self, x, t, name="") -> CoeffTriple:
        self.rows.append(np.array(x, dtype=np.float64))
        self._node(t)
        if self.calls != self.K or len(self.rows) != self.K:
            raise ValueError(f"sampler made {self.calls} model calls, expected {self.K}")
        M = np.stack(self.rows)
        A, B = M[:, : self.K], M[:, self.K:]
        # columns of noise never drawn stay zero; lower-triangular structure is checked by CoeffTriple
        return CoeffTriple(A, B, np.array(self.nodes), name=name)


def quadratic_time_grid(K: int, t_T=1.0, t_0=1e-3):
    """`time_quadratic` spacing (deps/dpm_solver_pytorch.py:475-478): the grid of weights/step_*_weight_*.npz"""
    return np.linspace(t_T ** 0.5, t_0 ** 0.5, K + 1) ** 2


def dpm_solver_pp_2s_triple(steps: int, t_T=1.0, t_0=1e-3, r1=0.5) -> CoeffTriple:
    """Singlestep DPM-Solver++(2S) on a uniform time grid, 2 model calls per step -> K = 2*steps rows
    (what src/AnalyzeDPMSolver.py:329-425 derives with sympy; update of deps/dpm_solver_pytorch.py:594-676)."""
    ns = VPLinearSchedule()
    ts = np.linspace(t_T, t_0, steps + 1)
    tr = CoefficientTracer(2 * steps, ns)
    x = tr.noise()
    for i in range(steps):
        s, t = ts[i], ts[i + 1]
        h = ns.lam(t) - ns.lam(s)
        s1 = ns.inv_lam(ns.lam(s) + r1 * h)
        y_s = tr.model_x0(x, s)
        x_s1 = ns.sigma(s1) / ns.sigma(s) * x - ns.alpha(s1) * np.expm1(-r1 * h) * y_s
        y_s1 = tr.model_x0(x_s1, s1)
        phi = ns.alpha(t) * np.expm1(-h)
        x = ns.sigma(t) / ns.sigma(s) * x - phi * y_s - (0.5 / r1) * phi * (y_s1 - y_s)
    return tr.finish(x, ts[-1], name=f"dpmsolverpp2s_{2 * steps:03d}")


def dpm_solver_2s_triple(steps: int, t_T=1.0, t_0=1e-3, r1=0.5) -> CoeffTriple:
    """Singlestep DPM-Solver-2 (noise prediction) on a uniform time grid (src/AnalyzeDPMSolver.py:228-326)."""
    ns = VPLinearSchedule()
    ts = np.linspace(t_T, t_0, steps + 1)
    tr = CoefficientTracer(2 * steps, ns)
    x = tr.noise()
    for i in range(steps):
        s, t = ts[i], ts[i + 1]
        h = ns.lam(t) - ns.lam(s)
        s1 = ns.inv_lam(ns.lam(s) + r1 * h)
        e_s = tr.model_eps(x, s)
        x_s1 = np.exp(ns.log_alpha(s1) - ns.log_alpha(s)) * x - ns.sigma(s1) * np.expm1(r1 * h) * e_s
        e_s1 = tr.model_eps(x_s1, s1)
        phi = ns.sigma(t) * np.expm1(h)
        x = np.exp(ns.log_alpha(t) - ns.log_alpha(s)) * x - phi * e_s - (0.5 / r1) * phi * (e_s1 - e_s)
    return tr.finish(x, ts[-1], name=f"dpmsolver2s_{2 * steps:03d}")


def dpm_solver_pp_2m_triple(K: int, grid="time_quadratic", t_T=1.0, t_0=1e-3) -> CoeffTriple:
    """Multistep DPM-Solver++(2M), one model call per step: first step first-order
    (deps/dpm_solver_pytorch.py:547-576), then the second-order multistep update (:796-831).  With the quadratic
    grid this is the sampler BASELINE config 3 names; the shipped step_15_weight_173 is a hand-tuned banded matrix on
    the same grid, this is the true solver's (dense) matrix."""
    ns = VPLinearSchedule()
    ts = quadratic_time_grid(K, t_T, t_0) if grid == "time_quadratic" else np.linspace(t_T, t_0, K + 1)
    tr = CoefficientTracer(K, ns)
    x = tr.noise()
    prev_y, prev_t = None, None
    for i in range(K):
        s, t = ts[i], ts[i + 1]
        y = tr.model_x0(x, s)
        h = ns.lam(t) - ns.lam(s)
        phi = ns.alpha(t) * np.expm1(-h)
        nxt = ns.sigma(t) / ns.sigma(s) * x - phi * y
        if prev_y is not None:
            r0 = (ns.lam(s) - ns.lam(prev_t)) / h
            nxt = nxt - 0.5 * phi * (1.0 / r0) * (y - prev_y)
        prev_y, prev_t, x = y, s, nxt
    return tr.finish(x, ts[-1], name=f"dpmsolverpp2m_{K:03d}")


def deis_tab_triple(K: int, ab_order: int = 3, t_T=1.0, t_0=1e-3, quad_points: int = 10000) -> CoeffTriple:
    """DEIS tAB-`ab_order` (exponential integrator, Adams-Bashforth in t) on the quadratic time grid, VP linear schedule:
        x_{i+1} = psi(t_i, t_{i+1}) x_i + sum_j C_ij eps(x_{i-j}, t_{i-j}),  order min(i, ab_order) at step i,
        C_ij = int_{t_i}^{t_{i+1}} psi(tau, t_{i+1}) * (-1/2 dlog(abar)/dtau / sqrt(1 - abar(tau))) * L_j(tau) dtau
    with L_j the Lagrange basis on (t_{i-o}..t_i) and the integral a left Riemann sum of `quad_points` points, exactly as
    deps/th_deis/multistep.py:6-96 + vpsde.py:39-63 evaluate it (there in jax float32; here float64).  The matrix the
    reference derives with sympy + jax in src/AnalyzeDEIS.py:90-138 (results/deis/deis_tab_*.npz)."""
    b0, b1 = 0.1, 20.0
    ns = VPLinearSchedule(b0, b1)
    ts = quadratic_time_grid(K, t_T, t_0)
    log_abar = lambda t: 2.0 * ns.log_alpha(t)
    abar = lambda t: np.exp(log_abar(t))
    dlog = lambda t: -t * (b1 - b0) - b0
    tr = CoefficientTracer(K, ns)
    x = tr.noise()
    eps_hist = []  # newest first
    for i in range(K):
        s, t = ts[i], ts[i + 1]
        o = min(i, ab_order)
        eps_hist.insert(0, tr.model_eps(x, s))
        tau = np.linspace(s, t, quad_points, endpoint=False)
        dt = (t - s) / quad_points
        integrand = np.sqrt(abar(t) / abar(tau)) * (-0.5 * dlog(tau) / np.sqrt(1.0 - abar(tau)))
        nodes = ts[i - o: i + 1]  # t_{i-o} .. t_i
        nxt = np.sqrt(abar(t) / abar(s)) * x
        for j in range(o + 1):
            idx = o - j  # node of eps_{i-j}
            num = tau[:, None] - nodes[None, :]
            den = nodes[idx] - nodes
            num[:, idx], den[idx] = 1.0, 1.0
            poly = np.prod(num, axis=1) / np.prod(den)
            nxt = nxt + float(np.sum(integrand * poly) * dt) * eps_hist[j]
        eps_hist = eps_hist[:ab_order]
        x = nxt
    return tr.finish(x, ts[-1], name=f"deis_tab_{K:03d}")


def _vp_euler_grid(num_step: int):
    n = num_step + 1
    return 1.0 + np.arange(n) * (1.0 / n - 1.0) / (n - 1), (1.0 / n - 1.0) / (n - 1)


def vp_euler_triple(num_step: int, kind: str = "ode") -> CoeffTriple:
    """Euler discretisations of the VP SDE on the uniform grid t: 1 -> 1/(K+1) with the score written through the
    x0-prediction, score = (alpha*y - x)/sigma^2 (src/AnalyzeEulerHeun.py:50-123 probability-flow ODE, :125-201
    Euler-Maruyama reverse SDE, :203-290 Heun).  kind: "ode" | "sde" | "heun" (2 model calls per step -> 2K rows).
    "heun" keeps the reference's second-stage quirk (alpha of the START node multiplies the second prediction,
    :249) so that the shipped ode_heun_* matrices are reproduced; pass kind="heun_exact" for the textbook update."""
    ns = VPLinearSchedule()
    ts, dt = _vp_euler_grid(num_step)
    beta = lambda t: ns.b0 + t * (ns.b1 - ns.b0)
    calls = 2 * num_step if kind.startswith("heun") else num_step
    tr = CoefficientTracer(calls, ns)
    x = tr.noise()

    def velocity(xv, y, t, alpha_t, half):
        score = (alpha_t * y - xv) / ns.sigma(t) ** 2
        return -0.5 * beta(t) * xv - (0.5 if half else 1.0) * beta(t) * score

    for i in range(num_step):
        s, t = ts[i], ts[i + 1]
        y_s = tr.model_x0(x, s)
        if kind == "ode":
            x = x + velocity(x, y_s, s, ns.alpha(s), True) * dt
        elif kind == "sde":
            x = x + velocity(x, y_s, s, ns.alpha(s), False) * dt + np.sqrt(beta(s)) * np.sqrt(abs(dt)) * tr.noise()
        else:
            v_s = velocity(x, y_s, s, ns.alpha(s), True)
            x_hat = x + v_s * dt
            y_hat = tr.model_x0(x_hat, t + 0.0005)  # the reference tags the predictor node with a tiny time offset
            a2 = ns.alpha(s) if kind == "heun" else ns.alpha(t)
            score_t = (a2 * y_hat - x_hat) / ns.sigma(t) ** 2
            v_t = -0.5 * beta(t) * x_hat - 0.5 * beta(t) * score_t
            x = x + 0.5 * (v_s + v_t) * dt
    name = {"ode": "ode_euler", "sde": "sde_euler"}.get(kind, "ode_heun")
    return tr.finish(x, ts[-1], name=f"{name}_{calls:03d}")


def dpm_solver_3s_triple(steps: int, plus_plus: bool = False, t_T=1.0, t_0=1e-3) -> CoeffTriple:
    """Singlestep third-order DPM-Solver-3 (noise prediction) / DPM-Solver++(3S) (data prediction) on a uniform time
    grid, r1 = 1/3, r2 = 2/3, 3 model calls per step -> K = 3*steps rows, with the update formulas exactly as the
    reference analyses them (src/AnalyzeDPMSolver.py:431-547 and :550-695 -- note the ++ variant there subtracts the
    difference terms; the shipped results/dpmsolverpp/dpmsolverpp3s_* matrices are reproduced as they are)."""
    ns = VPLinearSchedule()
    ts = np.linspace(t_T, t_0, steps + 1)
    r1, r2 = 1.0 / 3.0, 2.0 / 3.0
    tr = CoefficientTracer(3 * steps, ns)
    x = tr.noise()
    for i in range(steps):
        s, t = ts[i], ts[i + 1]
        h = ns.lam(t) - ns.lam(s)
        s1, s2 = ns.inv_lam(ns.lam(s) + r1 * h), ns.inv_lam(ns.lam(s) + r2 * h)
        if plus_plus:
            m_s = tr.model_x0(x, s)
            x_s1 = ns.sigma(s1) / ns.sigma(s) * x - ns.alpha(s1) * np.expm1(-r1 * h) * m_s
            m_s1 = tr.model_x0(x_s1, s1)
            x_s2 = (ns.sigma(s2) / ns.sigma(s) * x - ns.alpha(s2) * np.expm1(-r2 * h) * m_s
                    - (r2 / r1) * ns.alpha(s2) * (np.expm1(-r2 * h) / (r2 * h) + 1.0) * (m_s1 - m_s))
            m_s2 = tr.model_x0(x_s2, s2)
            x = (ns.sigma(t) / ns.sigma(s) * x - ns.alpha(t) * np.expm1(-h) * m_s
                 - (1.0 / r2) * ns.alpha(t) * (np.expm1(-h) / h + 1.0) * (m_s2 - m_s))
        else:
            e_s = tr.model_eps(x, s)
            x_s1 = np.exp(ns.log_alpha(s1) - ns.log_alpha(s)) * x - ns.sigma(s1) * np.expm1(r1 * h) * e_s
            e_s1 = tr.model_eps(x_s1, s1)
            x_s2 = (np.exp(ns.log_alpha(s2) - ns.log_alpha(s)) * x - ns.sigma(s2) * np.expm1(r2 * h) * e_s
                    - (r2 / r1) * ns.sigma(s2) * (np.expm1(r2 * h) / (r2 * h) - 1.0) * (e_s1 - e_s))
            e_s2 = tr.model_eps(x_s2, s2)
            x = (np.exp(ns.log_alpha(t) - ns.log_alpha(s)) * x - ns.sigma(t) * np.expm1(h) * e_s
                 - (1.0 / r2) * ns.sigma(t) * (np.expm1(h) / h - 1.0) * (e_s2 - e_s))
    return tr.finish(x, ts[-1], name=("dpmsolverpp3s" if plus_plus else "dpmsolver3s") + f"_{3 * steps:03d}")


# ----------------------------------------------------------------------------------------------
# the reference's optimised matrices (weights/step_{5,10,15}_weight_*.npz): relative patterns
# ----------------------------------------------------------------------------------------------
def vp_quadratic_node(K: int, t_T=1.0, t_0=1e-3, schedule: "VPLinearSchedule" = None) -> np.ndarray:
    """node_coeff (K+1, 3) = (t, alpha, sigma) on the quadratic time grid with VP-linear marginals: what the shipped
    weights/step_*.npz carry (theirs to float32 round-off: t to 1e-7, sigma to 5e-8; alpha exactly)."""
    ns = schedule or VPLinearSchedule()
    t = quadratic_time_grid(K, t_T, t_0)
    return np.stack([t, [ns.alpha(v) for v in t], [ns.sigma(v) for v in t]], axis=1)


def relative_patterns(triple: CoeffTriple, decimals=2):
    """Each row of A relative to its diagonal, rounded: the form the optimised matrices were authored in (e.g. row 5 of
    step_15_weight_173 is [0.30, -0.14, 0.56, -0.77, 1] over its band).  Returns K arrays of length k+1."""
    out = []
    for k in range(triple.K):
        r = triple.A[k, : k + 1] / triple.A[k, k]
        out.append(r if decimals is None else np.round(r, decimals))
    return out


def relative_pattern_triple(patterns, node, name="") -> CoeffTriple:
    """Deterministic Natural Inference matrix from per-row relative patterns: row k = alpha_{k+1} * pattern_k / sum(pattern_k)
    (so that rowsum(A) = alpha, the signal coefficient of the marginal) and B[k,0] = sigma_{k+1} on the initial noise --
    the invariants of weights/step_*.npz.  pattern_k may be shorter than k+1: it is right-aligned
    on the diagonal (a band)."""
    node = np.asarray(node, dtype=np.float64)
    K = len(patterns)
    if node.shape != (K + 1, 3):
        raise ValueError(f"node must be ({K + 1}, 3) for {K} pattern rows")
    A, B = np.zeros((K, K)), np.zeros((K, K + 1))
    for k, pat in enumerate(patterns):
        pat = np.asarray(pat, dtype=np.float64)
        if pat.ndim != 1 or not 1 <= len(pat) <= k + 1:
            raise ValueError(f"pattern {k} must have between 1 and {k + 1} entries")
        if pat.sum() == 0:
            raise ValueError(f"pattern {k} sums to zero")
        A[k, k + 1 - len(pat): k + 1] = node[k + 1, 1] * pat / pat.sum()
        B[k, 0] = node[k + 1, 2]
    return CoeffTriple(A, B, node, name=name)
