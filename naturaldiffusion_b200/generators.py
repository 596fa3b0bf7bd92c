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
