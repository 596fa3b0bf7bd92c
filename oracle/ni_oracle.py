"""CPU oracle for the Natural Inference (NI) sampling step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``naturaldiffusion_b200/`` imports this
module; it is imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the reported CPU baseline -- never as the product path.

What it is: a restatement, in plain torch-on-CPU / numpy, of the arithmetic the
reference (blairstar/NaturalDiffusion) performs on the NI hot path, keeping the
reference's dtypes and rounding points (fp64 history for the CIFAR loop, fp32
product / fp64 accumulate for the DiT loop, storage-dtype accumulate for the SD3
loop).  Each function cites the reference file:line it follows.

Pinning: ``tests/golden/make_golden.py`` imports the reference's *own* functions
from /root/reference (third-party modules stubbed, see ``oracle/ref_loader.py``),
runs them on seeded inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
and against the shipped coefficient matrices (44 under results/).  Parity is therefore pinned.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# coefficient triples (reference: src/Utils.py:49 writer, read by position at
# src/CIFAR10NaturalInference.py:273 and src/ValidateNaturalInference.py:319)
# --------------------------------------------------------------------------------------


def load_triple(path):
    """(A, B, node) by POSITION, as ``np.load(p).values()`` does in the reference."""
    with np.load(path) as z:
        A, B, node = [z[k] for k in z.files]
    return A, B, node


def load_sd3_csv(path):
    """28x28 weight table; first column is the index, header row the sigmas.

    Follows ``pd.read_csv(path, index_col=0).to_numpy()`` (src/SD3NaturalInference.py:196)
    without pandas so the oracle has no dependency the product does not have.
    """
    # numpy's own csv reader -- deliberately NOT the line-splitting parser of the product (naturaldiffusion_b200/coeffs.py),
    # so that the checker and the thing checked do not share code for reading this table
    table = np.genfromtxt(path, delimiter=",", skip_header=1, dtype=np.float64)
    return np.atleast_2d(table)[:, 1:].copy()


def sd3_sigmas(num_step=28, shift=3.0, num_train=1000):
    """FlowMatchEulerDiscreteScheduler.set_timesteps(28) of diffusers (third party, absent
    here; version unpinned in the reference's requirements.txt:14), as called at
    src/SD3NaturalInference.py:188-190.  Published algorithm restated:
      __init__ : sigma_min = shift*s/(1+(shift-1)*s) at s = 1/num_train, sigma_max = 1
      set_timesteps(n): u = linspace(sigma_max, sigma_min, n); sigma = shift*u/(1+(shift-1)*u);
                        append a trailing 0; stored as float32.
    Anchored on the reference's own csv headers (sigma rounded to 2 decimals) and on the
    csv body (W[k,j] = round(100*(sigma_j - sigma_{j+1}), 2)), see tests/test_oracle_golden.py.
    """
    s_min = 1.0 / num_train
    sigma_min = shift * s_min / (1 + (shift - 1) * s_min)
    u = np.linspace(1.0, sigma_min, num_step, dtype=np.float32).astype(np.float64)
    sig = shift * u / (1 + (shift - 1) * u)
    return np.append(sig, 0.0).astype(np.float32)


def sd3_triple(W, sigmas):
    """csv table -> (A, B, node) of the common NI form:
    model input of step k+1 = sigma_{k+1}*noise + (1-sigma_{k+1}) * sum_j W[k,j] x0_j / sum_j W[k,j]
    (src/SD3NaturalInference.py:207-209 with :157-168)."""
    K = W.shape[0]
    sig = np.asarray(sigmas, dtype=np.float64)
    A = np.zeros((K, K))
    B = np.zeros((K, K + 1))
    for k in range(K):
        row = W[k, : k + 1]
        A[k, : k + 1] = (1.0 - sig[k + 1]) * row / row.sum()
        B[k, 0] = sig[k + 1]
    node = np.stack([sig, 1.0 - sig, sig], axis=1)
    return A, B, node


# --------------------------------------------------------------------------------------
# DDPM / DDIM schedule tables (reference: src/ValidateNaturalInference.py:28-174, same
# code duplicated at src/AnalyzeDDPMDDIM.py:20-123)
# --------------------------------------------------------------------------------------


def spaced_steps(num_timesteps: int, count: int):
    """Single-section case of ``space_timesteps(1000, str(count))``
    (src/ValidateNaturalInference.py:57-78): accumulate a float stride and
    round-half-even each node."""
    if count <= 1:
        stride = 1.0
    else:
        stride = (num_timesteps - 1) / (count - 1)
    cur, out = 0.0, []
    for _ in range(count):
        out.append(round(cur))
        cur += stride
    return sorted(set(out))


def _alphas_bar():
    betas = np.linspace(0.0001, 0.02, 1000, dtype=np.float64)
    return np.cumprod(1.0 - betas)


def skip_tables(num_step: int):
    """Quantities of skip_ddpm_coeff / skip_ddim_coeff on the sub-sampled grid
    (src/ValidateNaturalInference.py:98-174). Index 0 = lowest noise level."""
    idx = spaced_steps(1000, num_step)
    ab = _alphas_bar()[idx]
    ab_prev = np.append(1.0, ab[:-1])
    alphas = ab / ab_prev
    betas = 1.0 - alphas
    var = betas * (1.0 - ab_prev) / (1.0 - ab)
    out = dict(
        idx=idx,
        alphas_bar=ab,
        alphas=alphas,
        log_var=np.log(np.append(1e-5, var[1:])),
        xt2x0=np.sqrt(1.0 / ab),
        eps2x0=np.sqrt(1.0 / ab - 1.0),
        ddpm_x0=np.sqrt(ab_prev) * betas / (1.0 - ab),
        ddpm_xt=np.sqrt(alphas) * (1.0 - ab_prev) / (1.0 - ab),
    )
    rect = np.sqrt((1.0 - ab_prev) / (1.0 - ab))
    out["ddim_xt"] = rect
    out["ddim_x0"] = np.sqrt(ab_prev) - rect * np.sqrt(ab)
    return out


def _first_order_rows(coef_xt, coef_x0, std):
    """Unroll x_{s-1} = coef_xt[s] x_s + coef_x0[s] x0_s + std[s] eps into matrix rows
    (closed forms of src/AnalyzeDDPMDDIM.py:126-174, :297-340 and
    src/AnalyzeFlowMatching.py:20-59).  Sampling order runs from index K-1 down to 0."""
    K = len(coef_xt)
    A = np.zeros((K, K))
    B = np.zeros((K, K + 1))
    for start in range(K):
        r = K - start - 1  # row = state after the step that lands on level `start`
        eps = [np.prod(coef_xt[start:K])]
        xz = []
        for ii in range(start, K)[::-1]:
            f = float(np.prod(coef_xt[start:ii]))
            if std is not None:
                eps.append(float(std[ii]) * f)
            xz.append(float(coef_x0[ii]) * f)
        B[r, : len(eps)] = eps
        A[r, : len(xz)] = xz
    return A, B


def ddim_triple(num_step: int):
    """src/AnalyzeDDPMDDIM.py:297-340 (`ddim_analyze_coeff`)."""
    t = skip_tables(num_step)
    K = num_step
    node = np.zeros((K + 1, 3))
    node[0] = [999, 0.0, 1.0]
    for start in range(K):
        if start == 0:
            node[K - start] = [-1, 1.0, 0.0]
        else:
            node[K - start] = [t["idx"][start - 1], np.sqrt(t["alphas_bar"][start - 1]), np.sqrt(1 - t["alphas_bar"][start - 1])]
    A, B = _first_order_rows(t["ddim_xt"], t["ddim_x0"], None)
    return A, B, node


def ddpm_triple(num_step: int):
    """src/AnalyzeDDPMDDIM.py:126-174 (`ddpm_analyze_coeff`)."""
    t = skip_tables(num_step)
    K = num_step
    std = np.sqrt(np.exp(t["log_var"]))
    node = np.zeros((K + 1, 3))
    node[0] = [999, 0.0, 1.0]
    for start in range(K):
        if start == 0:
            node[K - start] = [-1, 1.0, 0.0]
        else:
            node[K - start] = [t["idx"][start - 1], np.sqrt(t["alphas_bar"][start - 1]), np.sqrt(1 - t["alphas_bar"][start - 1])]
    A, B = _first_order_rows(t["ddpm_xt"], t["ddpm_x0"], std)
    return A, B, node


def flow_euler_triple(num_step: int):
    """src/AnalyzeFlowMatching.py:20-59 (`flow_analyze_coeff`)."""
    sig = np.linspace(0, 1, num_step + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        cxt = sig[:-1] / sig[1:]
    cx0 = 1.0 - cxt
    K = num_step
    node = np.zeros((K + 1, 3))
    node[0] = [1.0, 0.0, 1.0]
    for start in range(K):
        node[K - start] = [sig[start], 1 - sig[start], sig[start]]
    A, B = _first_order_rows(cxt, cx0, None)
    return A, B, node


# --------------------------------------------------------------------------------------
# the per-step functions, reference dtypes and rounding points kept
# --------------------------------------------------------------------------------------


@torch.no_grad()
def cifar_data_fn(score_fn, xt, t, x_coeff, eps_coeff):
    """src/CIFAR10NaturalInference.py:219-230: pred_x0 = (score*sigma^2 + xt)/alpha in fp64."""
    vec_t = t * torch.ones(xt.shape[0])
    score = score_fn(xt, vec_t)
    xt64 = xt.to(torch.float64)
    s64 = score.to(torch.float64)
    e = torch.tensor(eps_coeff, dtype=torch.float64)
    a = torch.tensor(x_coeff, dtype=torch.float64)
    return (s64 * e**2 + xt64) / a


def cifar_weighted_sum(row, seq_x0):
    """src/CIFAR10NaturalInference.py:233-238: accumulate in the history dtype (fp64), cast fp32.
    Zero coefficients are multiplied too."""
    out = torch.zeros_like(seq_x0[0])
    for ii, x0 in enumerate(seq_x0):
        out += x0 * row[ii]
    return out.to(torch.float32)


def validate_weighted_sum(weights, seq_elem):
    """src/ValidateNaturalInference.py:198-204: fp32 product, fp64 accumulator, fp32 result."""
    out = torch.zeros_like(seq_elem[0]).to(torch.float64)
    for ii, elem in enumerate(seq_elem):
        out += elem * weights[ii]
    return out.to(torch.float32)


def sd3_weighted_sum(seq_xstarts, weights=None):
    """src/SD3NaturalInference.py:157-168: row len(seq)-1, accumulate in the tensors' dtype,
    divide by the python-float row sum."""
    n = len(seq_xstarts)
    acc_w = 0
    acc = torch.zeros_like(seq_xstarts[0])
    for ii, arr in enumerate(seq_xstarts):
        w = 1 if weights is None else weights[n - 1][ii]
        acc += arr * w
        acc_w += w
    return acc / acc_w


def euler_weighted_sum(seq_xstarts, cliplen=0):
    """src/SD3NaturalInference.py:61-69."""
    acc = torch.zeros_like(seq_xstarts[0][1])
    acc_w = 0
    for w, x in seq_xstarts[-cliplen:]:
        acc += w * x
        acc_w += w
    return acc, acc / acc_w


def vp_marginal_std(t, beta_0=0.1, beta_1=20.0):
    """deps/score_sde_pytorch/sde_lib.py:141-145, fp32 torch arithmetic on a [B] tensor."""
    lmc = -0.25 * t**2 * (beta_1 - beta_0) - 0.5 * t * beta_0
    return torch.sqrt(1.0 - torch.exp(2.0 * lmc))


def make_vp_score_fn(model_fn):
    """deps/score_sde_pytorch/models/utils.py:144-160 (continuous VP branch)."""

    def score_fn(x, t):
        h = model_fn(x, t * 999)
        std = vp_marginal_std(t)
        return -h / std[:, None, None, None]

    return score_fn


# --------------------------------------------------------------------------------------
# the three loops; each returns (final, trace) with trace = list of per-step dicts
# --------------------------------------------------------------------------------------


@torch.no_grad()
def cifar_ni_loop(A, B, node, score_fn, noise):
    """src/CIFAR10NaturalInference.py:292-306."""
    ts = node[:, 0]
    K = ts.shape[0] - 1
    seq_x0, trace = [], []
    x = noise
    for kk in range(K):
        pred_x0 = cifar_data_fn(score_fn, x, ts[kk], node[kk, 1], node[kk, 2])
        seq_x0.append(pred_x0)
        next_x0 = cifar_weighted_sum(A[kk], seq_x0)
        next_eps = B[kk, 0] * noise
        x = next_x0 + next_eps
        trace.append(dict(x0=pred_x0, x_next=x))
    return x, trace


@torch.no_grad()
def validate_ni_loop(A, B, node, eps_model: Callable, noise, fresh_noise: Sequence[torch.Tensor], cfg_scale=4.0):
    """src/ValidateNaturalInference.py:343-366.  ``eps_model(z, timestep:int) -> (cond, uncond)``
    plays `forward_cfg` (:185-195) minus the fuse; ``fresh_noise[k]`` is the tensor
    `torch.randn_like` returns at step k (:359)."""
    K = B.shape[0]
    t = skip_tables(K)
    c1 = torch.from_numpy(t["xt2x0"]).to(torch.float32).flip(0)
    c2 = torch.from_numpy(t["eps2x0"]).to(torch.float32).flip(0)
    seq_x0, seq_eps, trace = [], [noise], []
    z = noise.clone()
    for kk in range(K):
        cond, uncond = eps_model(z, int(node[kk, 0]))
        fuse = uncond + cfg_scale * (cond - uncond)
        pred_x0 = c1[kk] * z - c2[kk] * fuse
        seq_x0.append(pred_x0)
        seq_eps.append(fresh_noise[kk])
        z = validate_weighted_sum(A[kk], seq_x0) + validate_weighted_sum(B[kk], seq_eps)
        trace.append(dict(x0=pred_x0, x_next=z))
    return z, trace


@torch.no_grad()
def ddpm_original_loop(num_step, eps_model, noise, fresh_noise, cfg_scale=4.0):
    """src/ValidateNaturalInference.py:235-250 (ancestral sampling, the comparator).
    fresh_noise[k] is consumed at the k-th executed step (sampling order)."""
    t = skip_tables(num_step)
    f32 = lambda a: torch.from_numpy(a).to(torch.float32)
    c1, c2, cxt, cx0, lv = f32(t["xt2x0"]), f32(t["eps2x0"]), f32(t["ddpm_xt"]), f32(t["ddpm_x0"]), f32(t["log_var"])
    z = noise.clone()
    trace = []
    for n, ii in enumerate(range(num_step)[::-1]):
        cond, uncond = eps_model(z, t["idx"][ii])
        fuse = uncond + cfg_scale * (cond - uncond)
        x0 = c1[ii] * z - c2[ii] * fuse
        mean = cxt[ii] * z + cx0[ii] * x0
        z = mean + torch.exp(0.5 * lv[ii]) * fresh_noise[n]
        trace.append(dict(x0=x0, x_next=z))
    return z, trace


@torch.no_grad()
def ddim_original_loop(num_step, eps_model, noise, cfg_scale=4.0):
    """src/ValidateNaturalInference.py:288-302."""
    t = skip_tables(num_step)
    f32 = lambda a: torch.from_numpy(a).to(torch.float32)
    c1, c2, cxt, cx0 = f32(t["xt2x0"]), f32(t["eps2x0"]), f32(t["ddim_xt"]), f32(t["ddim_x0"])
    z = noise.clone()
    trace = []
    for ii in range(num_step)[::-1]:
        cond, uncond = eps_model(z, t["idx"][ii])
        fuse = uncond + cfg_scale * (cond - uncond)
        x0 = c1[ii] * z - c2[ii] * fuse
        z = cxt[ii] * z + cx0[ii] * x0
        trace.append(dict(x0=x0, x_next=z))
    return z, trace


@torch.no_grad()
def sd3_ni_loop(W, sigmas, v_model: Callable, noises, cfg_scale=7):
    """src/SD3NaturalInference.py:198-223.  ``v_model(x_in, k) -> (v_text, v_null)``.
    Arithmetic runs in ``noises.dtype`` (the reference uses fp16); sigmas are fp32 0-d tensors."""
    K = W.shape[0]
    sig = torch.as_tensor(np.asarray(sigmas), dtype=torch.float32)
    seq, trace = [], []
    out = None
    for kk in range(K):
        sigma = sig[kk]
        curr = sd3_weighted_sum(seq, W) if len(seq) != 0 else torch.zeros_like(noises)
        x_in = sigma * noises + (1 - sigma) * curr
        v_text, v_null = v_model(x_in, kk)
        x0_null = x_in - sigma * v_null
        x0_text = x_in - sigma * v_text
        x0 = x0_null + cfg_scale * (x0_text - x0_null)
        seq.append(x0)
        out = sd3_weighted_sum(seq, W)
        trace.append(dict(x_in=x_in, x0=x0, out=out))
    return out, trace


@torch.no_grad()
def sd3_euler_original_loop(sigmas, v_model, noises, cfg_scale=7):
    """src/SD3NaturalInference.py:104-127 with is_vanilla_update=True (the Euler comparator)."""
    sig = torch.as_tensor(np.asarray(sigmas), dtype=torch.float32)
    x = noises.clone()
    for i in range(len(sig) - 1):
        v_text, v_null = v_model(x, i)
        fuse = v_null + cfg_scale * (v_text - v_null)
        x = x + (sig[i + 1] - sig[i]) * fuse
    return x


# --------------------------------------------------------------------------------------
# common-form loop in fp64: the yardstick for "one fused step"
# --------------------------------------------------------------------------------------


def ni_step_f64(a, b, x_in, outs, A_row, x0_hist, B_row, eps_hist):
    """x0 = a*x + sum_m b[m]*out[m];  x_next = sum_j A_row[j]*x0_j + sum_j B_row[j]*eps_j, all in fp64."""
    x0 = a * x_in.to(torch.float64)
    for bm, o in zip(b, outs):
        x0 = x0 + bm * o.to(torch.float64)
    hist = list(x0_hist) + [x0]
    nxt = torch.zeros_like(x0)
    for j, h in enumerate(hist):
        if A_row[j] != 0:
            nxt += A_row[j] * h.to(torch.float64)
    for j, e in enumerate(eps_hist):
        if j < len(B_row) and B_row[j] != 0:
            nxt += B_row[j] * e.to(torch.float64)
    return x0, nxt


def to_pixel_u8(x):
    """Output stage: inverse scaler (x+1)/2 (deps/score_sde_pytorch/datasets.py:32-38 for centered
    data) then src/CIFAR10NaturalInference.py:212-216: NCHW->NHWC, clip(x*255,0,255), truncating uint8."""
    y = (x + 1.0) / 2.0
    y = y.permute(0, 2, 3, 1).contiguous().numpy()
    return np.clip(y * 255, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------
# DPM-Solver++(2M), the original multistep sampler (comparator for generated matrices)
# --------------------------------------------------------------------------------------


def vp_linear(t, b0=0.1, b1=20.0):
    """(log alpha, alpha, sigma, lambda) of the continuous VP schedule, as NoiseScheduleVP('linear') computes them
    (deps/dpm_solver_pytorch.py:128-154), on fp32 torch scalars like the reference's solver."""
    t = torch.as_tensor(t, dtype=torch.float32)
    la = -0.25 * t**2 * (b1 - b0) - 0.5 * t * b0
    sig = torch.sqrt(1.0 - torch.exp(2.0 * la))
    return la, torch.exp(la), sig, la - torch.log(sig)


@torch.no_grad()
def dpmpp_2m_original_loop(ts, eps_model, noise):
    """DPM_Solver(algorithm_type='dpmsolver++').sample(method='multistep', order=2, lower_order_final=False) on the
    time grid `ts` (K+1 nodes): first step = dpm_solver_first_update (deps/dpm_solver_pytorch.py:547-576), then
    multistep_dpm_solver_second_update (:796-831, solver_type 'dpmsolver'); the data-prediction model is
    x0 = (x - sigma*eps)/alpha (:262-271 `data_prediction_fn` without thresholding)."""
    x = noise.clone()
    prev_m, prev_t = None, None
    for i in range(len(ts) - 1):
        s, t = float(ts[i]), float(ts[i + 1])
        _, a_s, sg_s, lam_s = vp_linear(s)
        _, a_t, sg_t, lam_t = vp_linear(t)
        m = (x - sg_s * eps_model(x, s)) / a_s
        h = lam_t - lam_s
        phi = torch.expm1(-h)
        nxt = (sg_t / sg_s) * x - (a_t * phi) * m
        if prev_m is not None:
            lam_p = vp_linear(prev_t)[3]
            r0 = (lam_s - lam_p) / h
            nxt = nxt - 0.5 * (a_t * phi) * ((1.0 / r0) * (m - prev_m))
        prev_m, prev_t, x = m, s, nxt
    return x


# --------------------------------------------------------------------------------------
# DEIS tAB3, the original multistep sampler (comparator for the generated matrix)
# --------------------------------------------------------------------------------------


def deis_tab_coefficients(ts, ab_order=3, num_item=10000, b0=0.1, b1=20.0):
    """[x_coef, C_0..C_ab_order] per step, as deps/th_deis/multistep.py:6-96 builds `ab_coef` (get_ab_eps_coef with the
    order ramp 0,1,..,ab_order; left Riemann sum with num_item points; VPSDE.psi / eps_integrand of
    deps/th_deis/vpsde.py:57-63 for the linear-beta alpha of :13-19), float64 numpy instead of jax float32."""
    ts = np.asarray(ts, dtype=np.float64)
    la = lambda t: 2.0 * (-0.25 * t**2 * (b1 - b0) - 0.5 * t * b0)
    coef = np.zeros((len(ts) - 1, ab_order + 2))
    for i in range(len(ts) - 1):
        t_start, t_end = ts[i], ts[i + 1]
        order = min(i, ab_order)
        coef[i, 0] = np.sqrt(np.exp(la(t_end) - la(t_start)))
        t_inter = np.linspace(t_start, t_end, num_item, endpoint=False)
        dt = (t_end - t_start) / num_item
        psi = np.sqrt(np.exp(la(t_end) - la(t_inter)))
        integrand = psi * (-0.5 * (-t_inter * (b1 - b0) - b0) / np.sqrt(1 - np.exp(la(t_inter))))
        ts_poly = ts[i - order: i + 1]
        for out_j, coef_idx in enumerate(range(order, -1, -1)):  # "we do flip of j here"
            poly = np.ones_like(t_inter)
            for k in range(order + 1):
                if k != coef_idx:
                    poly *= (t_inter - ts_poly[k]) / (ts_poly[coef_idx] - ts_poly[k])
            coef[i, 1 + out_j] = np.sum(integrand * poly) * dt
    return coef


@torch.no_grad()
def deis_tab_original_loop(ts, eps_model, noise, ab_order=3):
    """th_deis tAB sampler (deps/th_deis/multistep.py:98-104 `ab_step`, driven as in src/AnalyzeDEIS.py:42-58)."""
    coef = torch.from_numpy(deis_tab_coefficients(ts, ab_order)).to(noise.dtype)
    x = noise.clone()
    eps_pred = [x] * ab_order
    for i in range(len(ts) - 1):
        new_eps = eps_model(x, float(ts[i]))
        full = [new_eps, *eps_pred]
        nxt = coef[i, 0] * x
        for c, e in zip(coef[i, 1:], full):
            nxt = nxt + c * e
        x, eps_pred = nxt, full[:-1]
    return x


# --------------------------------------------------------------------------------------
# DPM-Solver / DPM-Solver++ `sample()` in general: multistep | singlestep, order 1..3 (comparators for the generated
# matrices of the samplers heading results/FID/dpmsolver*_*.csv).  Restated from deps/dpm_solver_pytorch.py with the
# reference's own control flow (model_prev_list / t_prev_list; inner time grids for r1, r2); pinned in
# tests/test_oracle_golden.py by running THIS code in coefficient space against matrices produced by the reference's
# unmodified DPM_Solver class (tests/golden/solver_matrices.npz).
# --------------------------------------------------------------------------------------
class _VPSchedule:
    """NoiseScheduleVP('linear') (deps/dpm_solver_pytorch.py:128-167) on scalars of `dtype`."""

    def __init__(self, dtype=torch.float32, b0=0.1, b1=20.0):
        self.dt, self.b0, self.b1 = dtype, b0, b1

    def T(self, v):
        return torch.as_tensor(v, dtype=self.dt)

    def log_alpha(self, t):
        t = self.T(t)
        return -0.25 * t**2 * (self.b1 - self.b0) - 0.5 * t * self.b0

    def alpha(self, t):
        return torch.exp(self.log_alpha(t))

    def std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * torch.log(1.0 - torch.exp(2.0 * la))

    def inverse_lambda(self, lamb):
        lamb = self.T(lamb)
        tmp = 2.0 * (self.b1 - self.b0) * torch.logaddexp(-2.0 * lamb, torch.zeros((1,), dtype=self.dt))[0]
        delta = self.b0**2 + tmp
        return tmp / (torch.sqrt(delta) + self.b0) / (self.b1 - self.b0)

    def time_steps(self, skip_type, t_T, t_0, N):
        """get_time_steps (:455-482)"""
        if skip_type == "time_uniform":
            return torch.linspace(float(t_T), float(t_0), N + 1, dtype=self.dt)
        if skip_type == "time_quadratic":
            return torch.linspace(float(t_T) ** 0.5, float(t_0) ** 0.5, N + 1, dtype=self.dt) ** 2
        if skip_type == "logSNR":
            return torch.stack([self.inverse_lambda(v) for v in torch.linspace(float(self.lam(t_T)), float(self.lam(t_0)), N + 1, dtype=self.dt)])
        raise ValueError(skip_type)


@torch.no_grad()
def dpm_solver_original_sample(eps_model, noise, steps, algorithm="dpmsolver++", method="multistep", order=3, skip_type="time_quadratic",
                               t_T=1.0, t_0=1e-3, lower_order_final=False, dtype=torch.float32):
    """`DPM_Solver(eps_model, NoiseScheduleVP('linear'), algorithm_type).sample(noise, steps, t_T, t_0, order, skip_type, method,
    lower_order_final, denoise_to_zero=False)` (:1166-1232), solver_type 'dpmsolver', no thresholding."""
    ns = _VPSchedule(dtype)
    pp = algorithm == "dpmsolver++"

    def model_fn(x, t):  # :262-271 data_prediction_fn / noise_prediction_fn
        e = eps_model(x, float(t))
        return (x - ns.std(t) * e) / ns.alpha(t) if pp else e

    def first(x, s, t, model_s):  # :547-592
        h = ns.lam(t) - ns.lam(s)
        if pp:
            return ns.std(t) / ns.std(s) * x - ns.alpha(t) * torch.expm1(-h) * model_s
        return torch.exp(ns.log_alpha(t) - ns.log_alpha(s)) * x - (ns.std(t) * torch.expm1(h)) * model_s

    def single2(x, s, t, r1):  # :594-676
        h = ns.lam(t) - ns.lam(s)
        s1 = ns.inverse_lambda(ns.lam(s) + r1 * h)
        model_s = model_fn(x, s)
        if pp:
            phi_11, phi_1 = torch.expm1(-r1 * h), torch.expm1(-h)
            x_s1 = (ns.std(s1) / ns.std(s)) * x - (ns.alpha(s1) * phi_11) * model_s
            model_s1 = model_fn(x_s1, s1)
            return (ns.std(t) / ns.std(s)) * x - (ns.alpha(t) * phi_1) * model_s - (0.5 / r1) * (ns.alpha(t) * phi_1) * (model_s1 - model_s)
        phi_11, phi_1 = torch.expm1(r1 * h), torch.expm1(h)
        x_s1 = torch.exp(ns.log_alpha(s1) - ns.log_alpha(s)) * x - (ns.std(s1) * phi_11) * model_s
        model_s1 = model_fn(x_s1, s1)
        return torch.exp(ns.log_alpha(t) - ns.log_alpha(s)) * x - (ns.std(t) * phi_1) * model_s - (0.5 / r1) * (ns.std(t) * phi_1) * (model_s1 - model_s)

    def single3(x, s, t, r1, r2):  # :677-795
        h = ns.lam(t) - ns.lam(s)
        s1, s2 = ns.inverse_lambda(ns.lam(s) + r1 * h), ns.inverse_lambda(ns.lam(s) + r2 * h)
        model_s = model_fn(x, s)
        if pp:
            phi_11, phi_12, phi_1 = torch.expm1(-r1 * h), torch.expm1(-r2 * h), torch.expm1(-h)
            phi_22, phi_2 = phi_12 / (r2 * h) + 1.0, phi_1 / h + 1.0
            x_s1 = (ns.std(s1) / ns.std(s)) * x - (ns.alpha(s1) * phi_11) * model_s
            model_s1 = model_fn(x_s1, s1)
            x_s2 = (ns.std(s2) / ns.std(s)) * x - (ns.alpha(s2) * phi_12) * model_s + r2 / r1 * (ns.alpha(s2) * phi_22) * (model_s1 - model_s)
            model_s2 = model_fn(x_s2, s2)
            return (ns.std(t) / ns.std(s)) * x - (ns.alpha(t) * phi_1) * model_s + (1.0 / r2) * (ns.alpha(t) * phi_2) * (model_s2 - model_s)
        phi_11, phi_12, phi_1 = torch.expm1(r1 * h), torch.expm1(r2 * h), torch.expm1(h)
        phi_22, phi_2 = phi_12 / (r2 * h) - 1.0, phi_1 / h - 1.0
        x_s1 = torch.exp(ns.log_alpha(s1) - ns.log_alpha(s)) * x - (ns.std(s1) * phi_11) * model_s
        model_s1 = model_fn(x_s1, s1)
        x_s2 = torch.exp(ns.log_alpha(s2) - ns.log_alpha(s)) * x - (ns.std(s2) * phi_12) * model_s - r2 / r1 * (ns.std(s2) * phi_22) * (model_s1 - model_s)
        model_s2 = model_fn(x_s2, s2)
        return torch.exp(ns.log_alpha(t) - ns.log_alpha(s)) * x - (ns.std(t) * phi_1) * model_s - (1.0 / r2) * (ns.std(t) * phi_2) * (model_s2 - model_s)

    def multi2(x, model_prev_list, t_prev_list, t):  # :796-852
        model_prev_1, model_prev_0 = model_prev_list[-2], model_prev_list[-1]
        t_prev_1, t_prev_0 = t_prev_list[-2], t_prev_list[-1]
        h_0, h = ns.lam(t_prev_0) - ns.lam(t_prev_1), ns.lam(t) - ns.lam(t_prev_0)
        D1_0 = (1.0 / (h_0 / h)) * (model_prev_0 - model_prev_1)
        if pp:
            phi_1 = torch.expm1(-h)
            return (ns.std(t) / ns.std(t_prev_0)) * x - (ns.alpha(t) * phi_1) * model_prev_0 - 0.5 * (ns.alpha(t) * phi_1) * D1_0
        phi_1 = torch.expm1(h)
        return torch.exp(ns.log_alpha(t) - ns.log_alpha(t_prev_0)) * x - (ns.std(t) * phi_1) * model_prev_0 - 0.5 * (ns.std(t) * phi_1) * D1_0

    def multi3(x, model_prev_list, t_prev_list, t):  # :854-904
        model_prev_2, model_prev_1, model_prev_0 = model_prev_list
        t_prev_2, t_prev_1, t_prev_0 = t_prev_list
        h_1, h_0, h = ns.lam(t_prev_1) - ns.lam(t_prev_2), ns.lam(t_prev_0) - ns.lam(t_prev_1), ns.lam(t) - ns.lam(t_prev_0)
        r0, r1 = h_0 / h, h_1 / h
        D1_0 = (1.0 / r0) * (model_prev_0 - model_prev_1)
        D1_1 = (1.0 / r1) * (model_prev_1 - model_prev_2)
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
        if pp:
            phi_1 = torch.expm1(-h)
            phi_2 = phi_1 / h + 1.0
            phi_3 = phi_2 / h - 0.5
            return (ns.std(t) / ns.std(t_prev_0)) * x - (ns.alpha(t) * phi_1) * model_prev_0 + (ns.alpha(t) * phi_2) * D1 - (ns.alpha(t) * phi_3) * D2
        phi_1 = torch.expm1(h)
        phi_2 = phi_1 / h - 1.0
        phi_3 = phi_2 / h - 0.5
        return (torch.exp(ns.log_alpha(t) - ns.log_alpha(t_prev_0)) * x - (ns.std(t) * phi_1) * model_prev_0 - (ns.std(t) * phi_2) * D1
                - (ns.std(t) * phi_3) * D2)

    def multistep_update(x, model_prev_list, t_prev_list, t, o):  # :932-954
        if o == 1:
            return first(x, t_prev_list[-1], t, model_prev_list[-1])
        return multi2(x, model_prev_list, t_prev_list, t) if o == 2 else multi3(x, model_prev_list, t_prev_list, t)

    x = noise.clone()
    if method == "multistep":  # :1171-1213
        timesteps = ns.time_steps(skip_type, t_T, t_0, steps)
        t = timesteps[0]
        t_prev_list, model_prev_list = [t], [model_fn(x, t)]
        for step in range(1, order):
            t = timesteps[step]
            x = multistep_update(x, model_prev_list, t_prev_list, t, step)
            t_prev_list.append(t)
            model_prev_list.append(model_fn(x, t))
        for step in range(order, steps + 1):
            t = timesteps[step]
            step_order = min(order, steps + 1 - step) if (lower_order_final and steps < 10) else order
            x = multistep_update(x, model_prev_list, t_prev_list, t, step_order)
            for i in range(order - 1):
                t_prev_list[i], model_prev_list[i] = t_prev_list[i + 1], model_prev_list[i + 1]
            t_prev_list[-1] = t
            if step < steps:
                model_prev_list[-1] = model_fn(x, t)
        return x
    # singlestep (:1214-1232 with the order schedule of :514-538)
    if order == 3:
        Kk = steps // 3 + 1
        orders = [3] * (Kk - 2) + [2, 1] if steps % 3 == 0 else ([3] * (Kk - 1) + [1] if steps % 3 == 1 else [3] * (Kk - 1) + [2])
    elif order == 2:
        orders = [2] * (steps // 2) if steps % 2 == 0 else [2] * (steps // 2) + [1]
    else:
        orders = [1] * steps
    cum = [0]
    for o in orders:
        cum.append(cum[-1] + o)
    if skip_type == "logSNR":  # "To reproduce the results in DPM-Solver paper" (:534-536): K outer nodes uniform in logSNR
        timesteps_outer = ns.time_steps(skip_type, t_T, t_0, len(orders))
    else:
        timesteps_outer = ns.time_steps(skip_type, t_T, t_0, steps)[torch.tensor(cum)]
    for step, o in enumerate(orders):
        s, t = timesteps_outer[step], timesteps_outer[step + 1]
        inner = ns.time_steps(skip_type, float(s), float(t), o)
        lambda_inner = torch.stack([ns.lam(v) for v in inner])
        h = lambda_inner[-1] - lambda_inner[0]
        if o == 1:
            x = first(x, s, t, model_fn(x, s))
        elif o == 2:
            x = single2(x, s, t, (lambda_inner[1] - lambda_inner[0]) / h)
        else:
            x = single3(x, s, t, (lambda_inner[1] - lambda_inner[0]) / h, (lambda_inner[2] - lambda_inner[0]) / h)
    return x


# --------------------------------------------------------------------------------------
# DEIS rho-AB, rho-RK and iPNDM (deps/th_deis/sampler.py:50-160, rk.py, multistep.py): original loops on tensors.
# th_deis is jax code and jax is not installed here: restated, and pinned in tests/test_solver_family.py by running THESE loops in
# coefficient space against matrices produced by the reference's own th_deis executed with a numpy-backed stand-in for jax
# (oracle/jax_numpy_shim.py, tests/golden/deis_matrices.npz: 38 settings).
# --------------------------------------------------------------------------------------
def _deis_abar(t, b0=0.1, b1=20.0):
    t = np.asarray(t, dtype=np.float64)
    return np.exp(2.0 * (-0.25 * t**2 * (b1 - b0) - 0.5 * t * b0))


def _deis_t_of_abar(a, b0=0.1, b1=20.0):
    c = np.log(np.asarray(a, dtype=np.float64)) / 2.0
    qa, qb = 0.25 * (b1 - b0), 0.5 * b0
    return (-qb + np.sqrt(qb**2 - 4 * qa * c)) / 2 / qa          # vpsde.py:8-10 quad_root


def deis_rev_ts(num_step, ts_order=2, ts_phase="t", t1=1.0, t0=1e-3):
    """get_rev_ts (deps/th_deis/sde.py:59-91), continuous-time SDE"""
    rho = lambda t: np.sqrt((1 - _deis_abar(t)) / _deis_abar(t))        # vpsde.py:65-67 with alpha_start = 1
    t_of_rho = lambda r: _deis_t_of_abar(1.0 / (r**2 + 1.0))           # :69-73
    if ts_phase == "t":
        return np.power(np.linspace(t1 ** (1.0 / ts_order), t0 ** (1.0 / ts_order), num_step + 1), ts_order)
    r0, r1 = rho(t0), rho(t1)
    if ts_phase == "log":
        return t_of_rho(np.exp(np.linspace(np.log(r1), np.log(r0), num_step + 1)))
    return t_of_rho(np.power(r1 ** (1.0 / ts_order) + np.linspace(0, num_step, num_step + 1) / num_step * (r0 ** (1.0 / ts_order) - r1 ** (1.0 / ts_order)), ts_order))


def deis_rho_ab_coefficients(rhos, ab_order=3, num_item=10000):
    """get_ab_eps_coef with the HelperSDE of sampler.py:104-109 (psi = 1, integrand = 1): plain Adams-Bashforth weights on the
    rho grid, by the same 10000-point left Riemann sums; coef[i, j] multiplies eps_{i-j}."""
    rhos = np.asarray(rhos, dtype=np.float64)
    coef = np.zeros((len(rhos) - 1, ab_order + 1))
    for i in range(len(rhos) - 1):
        order = min(i, ab_order)
        inter = np.linspace(rhos[i], rhos[i + 1], num_item, endpoint=False)
        d = (rhos[i + 1] - rhos[i]) / num_item
        poly_ts = rhos[i - order: i + 1]
        for out_j, coef_idx in enumerate(range(order, -1, -1)):
            poly = np.ones_like(inter)
            for k in range(order + 1):
                if k != coef_idx:
                    poly *= (inter - poly_ts[k]) / (poly_ts[coef_idx] - poly_ts[k])
            coef[i, out_j] = np.sum(poly) * d
    return coef


@torch.no_grad()
def deis_original_sample(eps_model, noise, num_step, method="rho_rk", ab_order=3, rk_method="3kutta", ts_phase="t", ts_order=2):
    """th_deis.get_sampler(...)(noise) for method rho_ab | rho_rk | ipndm (t_ab: deis_tab_original_loop above)."""
    f32 = lambda v: torch.tensor(float(v), dtype=noise.dtype, device=noise.device)  # th_deis hands torch the coefficients in the tensors' dtype
    if method == "ipndm":  # sampler.py:50-95
        ts = deis_rev_ts(num_step, 1, "t")
        lin = [[1.0, 0, 0, 0], [1.5, -0.5, 0, 0], [23 / 12.0, -16 / 12.0, 5 / 12.0, 0], [55 / 24.0, -59 / 24.0, 37 / 24.0, -9 / 24.0]]
        x, eps_pred = noise.clone(), [noise] * 3
        for i in range(num_step):
            a_cur, a_next = _deis_abar(ts[i]), _deis_abar(ts[i + 1])
            ddim = np.sqrt(1 - a_next) - np.sqrt(a_next / a_cur) * np.sqrt(1 - a_cur)
            full = [eps_model(x, float(ts[i])), *eps_pred]
            nxt = f32(np.sqrt(a_next / a_cur)) * x
            for c, e in zip(lin[min(i, 3)], full):
                nxt = nxt + f32(ddim * c) * e
            x, eps_pred = nxt, full[:-1]
        return x
    ts = deis_rev_ts(num_step, ts_order, ts_phase)
    abar = _deis_abar(ts)
    rhos = np.sqrt((1 - abar) / abar)
    t_of_rho = lambda r: float(_deis_t_of_abar(1.0 / (r**2 + 1.0)))
    v2x = lambda v, t: v / f32(np.sqrt(1.0 / _deis_abar(t)))      # vpsde.py:75-81
    v = f32(np.sqrt(1.0 / abar[0])) * noise
    if method == "rho_ab":  # sampler.py:98-133
        coef = deis_rho_ab_coefficients(rhos, ab_order)
        eps_pred = [noise] * ab_order
        for i in range(num_step):
            t_cur = t_of_rho(rhos[i])
            full = [eps_model(v2x(v, t_cur), t_cur), *eps_pred]
            nxt = v.clone()
            for c, e in zip(coef[i], full):
                nxt = nxt + f32(c) * e
            v, eps_pred = nxt, full[:-1]
        return v2x(v, ts[-1])
    if method == "rho_rk":  # sampler.py:136-160 + rk.py
        tabs = {"1euler": ([0.0], [[]], [1.0]), "2heun": ([0.0, 1.0], [[], [1.0]], [0.5, 0.5]),
                "3kutta": ([0.0, 0.5, 1.0], [[], [0.5], [-1.0, 2.0]], [1.0 / 6, 4.0 / 6, 1.0 / 6]),
                "3heun": ([0.0, 1.0 / 3, 2.0 / 3], [[], [1.0 / 3], [0.0, 2.0 / 3]], [0.25, 0.0, 0.75]),
                "4rk": ([0.0, 0.5, 0.5, 1.0], [[], [0.5], [0.0, 0.5], [0.0, 0.0, 1.0]], [1.0 / 6, 2.0 / 6, 2.0 / 6, 1.0 / 6])}
        c, a, b = tabs[rk_method]
        fn = lambda vv, rho: eps_model(v2x(vv, t_of_rho(rho)), t_of_rho(rho))
        for i in range(num_step):
            dt = rhos[i + 1] - rhos[i]
            ks = []
            for st in range(len(c)):
                vv = v
                for q in range(st):
                    if a[st][q] != 0.0:
                        vv = vv + f32(dt * a[st][q]) * ks[q]
                ks.append(fn(vv, rhos[i] + dt * c[st]))
            for bq, kq in zip(b, ks):
                if bq != 0.0:
                    v = v + f32(dt * bq) * kq
        return v2x(v, ts[-1])
    raise ValueError(method)
