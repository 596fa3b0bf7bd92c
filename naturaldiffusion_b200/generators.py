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
# DPM-Solver / DPM-Solver++ exactly as the reference drives them for results/FID/dpmsolver*_{5,10,15}step.csv
# (src/CIFAR10NaturalInference.py:363-393: multistep | singlestep, order 2 | 3, time_quadratic, lower_order_final=False)
# ----------------------------------------------------------------------------------------------
def solver_time_steps(skip_type: str, t_T: float, t_0: float, N: int, ns: "VPLinearSchedule" = None):
    """`DPM_Solver.get_time_steps` (deps/dpm_solver_pytorch.py:455-482), float64."""
    if skip_type == "time_uniform":
        return np.linspace(t_T, t_0, N + 1)
    if skip_type == "time_quadratic":
        return np.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1) ** 2
    if skip_type == "logSNR":
        ns = ns or VPLinearSchedule()
        return ns.inv_lam(np.linspace(ns.lam(t_T), ns.lam(t_0), N + 1))
    raise ValueError(f"unsupported skip_type {skip_type!r}")


def singlestep_orders(steps: int, order: int):
    """order of every outer step so that the model calls add up to `steps` (deps/dpm_solver_pytorch.py:514-533)"""
    if order == 3:
        K = steps // 3 + 1
        return [3] * (K - 2) + [2, 1] if steps % 3 == 0 else [3] * (K - 1) + ([1] if steps % 3 == 1 else [2])
    if order == 2:
        return [2] * (steps // 2) if steps % 2 == 0 else [2] * (steps // 2) + [1]
    if order == 1:
        return [1] * steps
    raise ValueError("order must be 1, 2 or 3")


class _DpmUpdates:
    """The six update formulas of deps/dpm_solver_pytorch.py (solver_type 'dpmsolver') on coefficient vectors.  `pp`
    selects data prediction (DPM-Solver++, model value = x0) or noise prediction (DPM-Solver, model value = eps)."""

    def __init__(self, ns, pp: bool, tracer: "CoefficientTracer"):
        self.ns, self.pp, self.tr = ns, pp, tracer

    def model(self, x, t):
        return self.tr.model_x0(x, t) if self.pp else self.tr.model_eps(x, t)

    def _lin(self, s, t):
        """(coefficient of x, coefficient scale of the model value, h) of a step s -> t"""
        ns = self.ns
        h = ns.lam(t) - ns.lam(s)
        if self.pp:
            return ns.sigma(t) / ns.sigma(s), ns.alpha(t), h
        return np.exp(ns.log_alpha(t) - ns.log_alpha(s)), ns.sigma(t), h

    def _phi1(self, h):
        return np.expm1(-h) if self.pp else np.expm1(h)

    def first(self, x, s, t, m_s):                                   # :547-592
        cx, cm, h = self._lin(s, t)
        return cx * x - cm * self._phi1(h) * m_s

    def single_second(self, x, s, t, r1):                            # :594-676
        ns = self.ns
        m_s = self.model(x, s)
        h = ns.lam(t) - ns.lam(s)
        s1 = ns.inv_lam(ns.lam(s) + r1 * h)
        x_s1 = self.first(x, s, s1, m_s)
        m_s1 = self.model(x_s1, s1)
        cx, cm, _ = self._lin(s, t)
        phi_1 = self._phi1(h)
        return cx * x - cm * phi_1 * m_s - (0.5 / r1) * cm * phi_1 * (m_s1 - m_s)

    def single_third(self, x, s, t, r1, r2):                         # :677-795
        ns = self.ns
        m_s = self.model(x, s)
        h = ns.lam(t) - ns.lam(s)
        s1, s2 = ns.inv_lam(ns.lam(s) + r1 * h), ns.inv_lam(ns.lam(s) + r2 * h)
        x_s1 = self.first(x, s, s1, m_s)
        m_s1 = self.model(x_s1, s1)
        cx2, cm2, _ = self._lin(s, s2)
        cx, cm, _ = self._lin(s, t)
        if self.pp:
            phi_12, phi_1 = np.expm1(-r2 * h), np.expm1(-h)
            phi_22, phi_2 = phi_12 / (r2 * h) + 1.0, phi_1 / h + 1.0
            x_s2 = cx2 * x - cm2 * phi_12 * m_s + (r2 / r1) * cm2 * phi_22 * (m_s1 - m_s)
            m_s2 = self.model(x_s2, s2)
            return cx * x - cm * phi_1 * m_s + (1.0 / r2) * cm * phi_2 * (m_s2 - m_s)
        phi_12, phi_1 = np.expm1(r2 * h), np.expm1(h)
        phi_22, phi_2 = phi_12 / (r2 * h) - 1.0, phi_1 / h - 1.0
        x_s2 = cx2 * x - cm2 * phi_12 * m_s - (r2 / r1) * cm2 * phi_22 * (m_s1 - m_s)
        m_s2 = self.model(x_s2, s2)
        return cx * x - cm * phi_1 * m_s - (1.0 / r2) * cm * phi_2 * (m_s2 - m_s)

    def multi_second(self, x, m_prev, t_prev, t):                    # :796-852
        ns = self.ns
        (m1, m0), (t1, t0) = m_prev[-2:], t_prev[-2:]
        cx, cm, h = self._lin(t0, t)
        r0 = (ns.lam(t0) - ns.lam(t1)) / h
        D1_0 = (1.0 / r0) * (m0 - m1)
        phi_1 = self._phi1(h)
        return cx * x - cm * phi_1 * m0 - 0.5 * cm * phi_1 * D1_0

    def multi_third(self, x, m_prev, t_prev, t):                     # :854-904
        ns = self.ns
        (m2, m1, m0), (t2, t1, t0) = m_prev[-3:], t_prev[-3:]
        cx, cm, h = self._lin(t0, t)
        r0, r1 = (ns.lam(t0) - ns.lam(t1)) / h, (ns.lam(t1) - ns.lam(t2)) / h
        D1_0, D1_1 = (1.0 / r0) * (m0 - m1), (1.0 / r1) * (m1 - m2)
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
        if self.pp:
            phi_1 = np.expm1(-h)
            phi_2 = phi_1 / h + 1.0
            phi_3 = phi_2 / h - 0.5
            return cx * x - cm * phi_1 * m0 + cm * phi_2 * D1 - cm * phi_3 * D2
        phi_1 = np.expm1(h)
        phi_2 = phi_1 / h - 1.0
        phi_3 = phi_2 / h - 0.5
        return cx * x - cm * phi_1 * m0 - cm * phi_2 * D1 - cm * phi_3 * D2


def dpm_solver_triple(K: int, algorithm: str = "dpmsolver++", method: str = "multistep", order: int = 3,
                      skip_type: str = "time_quadratic", t_T: float = 1.0, t_0: float = 1e-3, lower_order_final: bool = False) -> CoeffTriple:
    """`DPM_Solver.sample(steps=K, order, skip_type, method, lower_order_final, denoise_to_zero=False)` of
    deps/dpm_solver_pytorch.py:1166-1232 as a K-row coefficient matrix (K = model calls).  multistep: order-1 start, then
    order 2, then order 3 (:1171-1213); singlestep: the order schedule that uses up K calls (:514-538), r1/r2 from the inner
    time grid (:1222-1227).  These are the samplers heading results/FID/dpmsolver*_*step.csv."""
    if algorithm not in ("dpmsolver", "dpmsolver++"):
        raise ValueError("algorithm must be 'dpmsolver' or 'dpmsolver++'")
    ns = VPLinearSchedule()
    tr = CoefficientTracer(K, ns)
    U = _DpmUpdates(ns, algorithm == "dpmsolver++", tr)
    x = tr.noise()
    if method == "multistep":
        if K < order:
            raise ValueError("multistep needs steps >= order")
        ts = solver_time_steps(skip_type, t_T, t_0, K, ns)
        t_prev, m_prev = [ts[0]], [U.model(x, ts[0])]

        def update(x, t, o):
            if o == 1:
                return U.first(x, t_prev[-1], t, m_prev[-1])
            return U.multi_second(x, m_prev, t_prev, t) if o == 2 else U.multi_third(x, m_prev, t_prev, t)

        for step in range(1, order):
            x = update(x, ts[step], step)
            t_prev.append(ts[step])
            m_prev.append(U.model(x, ts[step]))
        for step in range(order, K + 1):
            o = min(order, K + 1 - step) if (lower_order_final and K < 10) else order
            x = update(x, ts[step], o)
            t_prev = t_prev[1:] + [ts[step]]
            if step < K:
                m_prev = m_prev[1:] + [U.model(x, ts[step])]
        t_end = ts[-1]
    elif method in ("singlestep", "singlestep_fixed"):
        if method == "singlestep":
            orders = singlestep_orders(K, order)
            idx = np.cumsum([0] + orders)
            outer = solver_time_steps(skip_type, t_T, t_0, K, ns)[idx] if skip_type != "logSNR" else solver_time_steps(skip_type, t_T, t_0, len(orders), ns)
        else:
            orders = [order] * (K // order)
            if K % order:
                raise ValueError("singlestep_fixed needs K divisible by order")
            outer = solver_time_steps(skip_type, t_T, t_0, len(orders), ns)
        for i, o in enumerate(orders):
            s, t = outer[i], outer[i + 1]
            inner = solver_time_steps(skip_type, s, t, o, ns)
            lam = np.array([ns.lam(v) for v in inner])
            h = lam[-1] - lam[0]
            if o == 1:
                x = U.first(x, s, t, U.model(x, s))
            elif o == 2:
                x = U.single_second(x, s, t, (lam[1] - lam[0]) / h)
            else:
                x = U.single_third(x, s, t, (lam[1] - lam[0]) / h, (lam[2] - lam[0]) / h)
        t_end = outer[-1]
    else:
        raise ValueError("method must be 'multistep', 'singlestep' or 'singlestep_fixed'")
    tag = {"dpmsolver": "dpmsolver", "dpmsolver++": "dpmsolverpp"}[algorithm]
    return tr.finish(x, t_end, name=f"{tag}_{method}{order}_{K:03d}")


# ----------------------------------------------------------------------------------------------
# DEIS as the reference drives it for results/FID/deis_{5,10,15}step.csv (src/CIFAR10NaturalInference.py:166-178:
# ts_phase "t" | "rho", ts_order 2, method t_ab | rho_ab | rho_rk, ab_order 2 | 3) plus iPNDM
# ----------------------------------------------------------------------------------------------
class _DeisVP:
    """VP-linear quantities of deps/th_deis/vpsde.py:11-78 in float64: abar(t), psi, the eps integrand, rho(t) and t(rho)."""

    def __init__(self, beta_0=0.1, beta_1=20.0):
        self.b0, self.b1 = beta_0, beta_1

    def log_abar(self, t):
        return 2.0 * (-0.25 * t ** 2 * (self.b1 - self.b0) - 0.5 * t * self.b0)

    def abar(self, t):
        return np.exp(self.log_abar(t))

    def t_of_abar(self, a):
        c = np.log(a) / 2.0
        qa, qb = 0.25 * (self.b1 - self.b0), 0.5 * self.b0
        return (-qb + np.sqrt(qb * qb - 4.0 * qa * c)) / (2.0 * qa)

    def psi(self, t0, t1):
        return np.sqrt(self.abar(t1) / self.abar(t0))

    def eps_integrand(self, t):
        return -0.5 * (-t * (self.b1 - self.b0) - self.b0) / np.sqrt(1.0 - self.abar(t))

    def rho(self, t):
        a = self.abar(t)
        return np.sqrt((1.0 - a) / a)

    def t_of_rho(self, rho):
        return self.t_of_abar(1.0 / (rho * rho + 1.0))

    def rev_ts(self, num_step, ts_order=2, ts_phase="t", t1=1.0, t0=1e-3):
        """deps/th_deis/sde.py:59-91 (continuous-time branch)"""
        if ts_phase == "t":
            return np.linspace(t1 ** (1.0 / ts_order), t0 ** (1.0 / ts_order), num_step + 1) ** ts_order
        r0, r1 = self.rho(t0), self.rho(t1)
        if ts_phase == "log":
            return self.t_of_rho(np.exp(np.linspace(np.log(r1), np.log(r0), num_step + 1)))
        if ts_phase == "rho":
            p = 1.0 / ts_order
            return self.t_of_rho((r1 ** p + np.arange(num_step + 1) / num_step * (r0 ** p - r1 ** p)) ** ts_order)
        raise ValueError("ts_phase must be 't', 'log' or 'rho'")


def _ab_coefficients(grid, psi, integrand, ab_order: int, quad_points: int = 10000):
    """`get_ab_eps_coef` (deps/th_deis/multistep.py:6-96): C[i, j] multiplies eps_{i-j}; step i uses order min(i, ab_order);
    C_ij = left Riemann sum over [g_i, g_{i+1}) of psi(tau, g_{i+1}) * integrand(tau) * Lagrange_j(tau)."""
    n = len(grid) - 1
    C = np.zeros((n, ab_order + 1))
    for i in range(n):
        s, t = grid[i], grid[i + 1]
        o = min(i, ab_order)
        tau = np.linspace(s, t, quad_points, endpoint=False)
        dt = (t - s) / quad_points
        w = psi(tau, t) * integrand(tau)
        nodes = grid[i - o: i + 1]
        for j in range(o + 1):
            k = o - j  # node of eps_{i-j}
            num = tau[:, None] - nodes[None, :]
            den = nodes[k] - nodes
            num[:, k], den[k] = 1.0, 1.0
            C[i, j] = float(np.sum(w * np.prod(num, axis=1) / np.prod(den)) * dt)
    return C


RK_TABLEAUS = {  # deps/th_deis/rk.py: (c, a rows, b)
    "1euler": ([0.0], [[]], [1.0]),
    "2heun": ([0.0, 1.0], [[], [1.0]], [0.5, 0.5]),
    "3kutta": ([0.0, 0.5, 1.0], [[], [0.5], [-1.0, 2.0]], [1 / 6, 4 / 6, 1 / 6]),
    "3ral": ([0.0, 0.5, 0.75], [[], [0.5], [0.0, 0.75]], [2 / 9, 1 / 3, 4 / 9]),
    "3heun": ([0.0, 1 / 3, 2 / 3], [[], [1 / 3], [0.0, 2 / 3]], [0.25, 0.0, 0.75]),
    "3vdh": ([0.0, 8 / 15, 2 / 3], [[], [8 / 15], [0.25, 5 / 12]], [0.25, 0.0, 0.75]),
    "3ssprk": ([0.0, 1.0, 0.5], [[], [1.0], [0.25, 0.25]], [1 / 6, 1 / 6, 2 / 3]),
    "4rk": ([0.0, 0.5, 0.5, 1.0], [[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]], [1 / 6, 2 / 6, 2 / 6, 1 / 6]),
}


def deis_triple(num_step: int, method: str = "rho_rk", ab_order: int = 3, rk_method: str = "3kutta", ts_phase: str = "t",
                ts_order: int = 2, t_T: float = 1.0, t_0: float = 1e-3, quad_points: int = 10000) -> CoeffTriple:
    """`th_deis.get_sampler(sde, eps_fn, ts_phase, ts_order, num_step, method, ab_order, rk_method)` (deps/th_deis/sampler.py:15-160)
    on the VP-linear SDE as a coefficient matrix:
      t_ab    exponential integrator, Adams-Bashforth in t (:26-49)                              K = num_step rows
      rho_ab  Adams-Bashforth in rho on v = x / sqrt(abar): dv/drho = eps (:98-133)               K = num_step rows
      rho_rk  Runge-Kutta in rho (default Kutta's third order, rk.py:17-25) (:136-160)            K = stages * num_step rows
      ipndm   DDIM step with the linear-multistep combination of the last 4 eps, uniform t grid (:50-95)
    The model is always called on x = v * sqrt(abar(t)) at t = t(rho), as `eps_fn_vrho` does."""
    vp = _DeisVP()
    ns = VPLinearSchedule(vp.b0, vp.b1)
    if method == "t_ab":
        return deis_tab_triple(num_step, ab_order, t_T, t_0, quad_points) if (ts_phase == "t" and ts_order == 2) else _deis_tab_on(vp.rev_ts(num_step, ts_order, ts_phase, t_T, t_0), ab_order, quad_points)
    if method == "ipndm":
        ts = vp.rev_ts(num_step, 1, "t", t_T, t_0)
        tr = CoefficientTracer(num_step, ns)
        x, hist = tr.noise(), []
        lin = [[1.0], [1.5, -0.5], [23 / 12, -16 / 12, 5 / 12], [55 / 24, -59 / 24, 37 / 24, -9 / 24]]
        for i in range(num_step):
            s, t = ts[i], ts[i + 1]
            hist.insert(0, tr.model_eps(x, s))
            ddim = np.sqrt(1 - vp.abar(t)) - np.sqrt(vp.abar(t) / vp.abar(s)) * np.sqrt(1 - vp.abar(s))
            x = vp.psi(s, t) * x + ddim * sum(c * e for c, e in zip(lin[min(i, 3)], hist))
            hist = hist[:4]
        return tr.finish(x, ts[-1], name=f"deis_ipndm_{num_step:03d}")
    ts = vp.rev_ts(num_step, ts_order, ts_phase, t_T, t_0)
    rhos = vp.rho(ts)
    to_x = lambda v, t: v * np.sqrt(vp.abar(t))
    if method == "rho_ab":
        C = _ab_coefficients(rhos, lambda a, b: np.ones_like(a), lambda a: np.ones_like(a), ab_order, quad_points)
        tr = CoefficientTracer(num_step, ns)
        v, hist = tr.noise() / np.sqrt(vp.abar(ts[0])), []
        for i in range(num_step):
            t_i = vp.t_of_rho(rhos[i])
            hist.insert(0, tr.model_eps(to_x(v, t_i), t_i))
            v = v + sum(C[i, j] * hist[j] for j in range(min(i, ab_order) + 1))
            hist = hist[:ab_order]
        return tr.finish(to_x(v, ts[-1]), ts[-1], name=f"deis_rho_ab{ab_order}_{num_step:03d}")
    if method == "rho_rk":
        c, a, b = RK_TABLEAUS[rk_method]
        tr = CoefficientTracer(len(c) * num_step, ns)
        v = tr.noise() / np.sqrt(vp.abar(ts[0]))
        for i in range(num_step):
            d = rhos[i + 1] - rhos[i]
            ks = []
            for st in range(len(c)):
                vv = v + d * sum(a[st][q] * ks[q] for q in range(st)) if st else v
                t_st = vp.t_of_rho(rhos[i] + d * c[st])
                ks.append(tr.model_eps(to_x(vv, t_st), t_st))
            v = v + d * sum(bq * kq for bq, kq in zip(b, ks))
        return tr.finish(to_x(v, ts[-1]), ts[-1], name=f"deis_rho_rk_{rk_method}_{len(c) * num_step:03d}")
    raise ValueError("method must be t_ab, rho_ab, rho_rk or ipndm")


def _deis_tab_on(ts, ab_order, quad_points):
    """t_ab on an arbitrary reverse time grid (deis_tab_triple is the quadratic-grid special case)"""
    vp = _DeisVP()
    ns = VPLinearSchedule(vp.b0, vp.b1)
    K = len(ts) - 1
    C = _ab_coefficients(ts, vp.psi, vp.eps_integrand, ab_order, quad_points)
    tr = CoefficientTracer(K, ns)
    x, hist = tr.noise(), []
    for i in range(K):
        hist.insert(0, tr.model_eps(x, ts[i]))
        x = vp.psi(ts[i], ts[i + 1]) * x + sum(C[i, j] * hist[j] for j in range(min(i, ab_order) + 1))
        hist = hist[:ab_order]
    return tr.finish(x, ts[-1], name=f"deis_tab_{K:03d}")


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
