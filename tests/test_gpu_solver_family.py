"""GPU: the original samplers of the reference's FID tables, expressed as generated coefficient matrices and run through
the fused step, against the oracle's restatements of the ORIGINAL sampler loops on the same noise.
DPM-Solver / DPM-Solver++ multistep-3 and singlestep-3 (deps/dpm_solver_pytorch.py:675-904), DEIS rho-RK (Kutta), rho-AB,
iPNDM (deps/th_deis/sampler.py:50-160).  Tolerance: max|d| <= 1e-5 * ||ref||_2 (north-star) -- except where the original
solver is itself ill-conditioned in fp32 (singlestep-3 at 5 steps has matrix entries of 90), stated per case."""
import numpy as np
import pytest
import torch

from naturaldiffusion_b200 import generators as G
from naturaldiffusion_b200.ops import philox_normal
from naturaldiffusion_b200.sampler import NaturalInferenceSampler
from oracle import ni_oracle as O
from toy_models import ToyEps

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"


def rel_err(got, ref):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    return ((got - ref).abs().max() / ref.norm().clamp_min(1e-30)).item()


def _run(triple, B, seed):
    ns = G.VPLinearSchedule()
    ts = triple.node[:, 0]
    io = [(1.0 / ns.alpha(t), -ns.sigma(t) / ns.alpha(t), 0.0) for t in ts[:-1]]  # x0 = (x - sigma*eps)/alpha
    net = ToyEps(3, seed=11, t_scale=0.3)
    eps_model = lambda x, t: float(ns.sigma(t)) * x + 0.1 * net(x, float(t))      # bounded eps-predictor
    s = NaturalInferenceSampler(triple, io, B, (3, 32, 32), device=DEV, seed=seed)
    x = s.sample(lambda z, k: eps_model(z, ts[k]))
    noise = philox_normal((B, 3, 32, 32), seed=seed, tensor_id=0, device=DEV)
    return x, noise, eps_model, s


@pytest.mark.parametrize("alg", ["dpmsolver", "dpmsolver++"])
@pytest.mark.parametrize("method,order,K,tol", [("multistep", 3, 10, 1e-5), ("multistep", 3, 15, 1e-5), ("multistep", 3, 5, 1e-5),
                                                ("singlestep", 3, 15, 1e-5), ("singlestep", 3, 10, 2e-5), ("singlestep", 2, 10, 1e-5)])
def test_dpm_solver_matrix_equals_original_solver(alg, method, order, K, tol):
    triple = G.dpm_solver_triple(K, alg, method, order)
    x, noise, eps_model, s = _run(triple, 128, seed=21)
    assert not s.plan.markov
    xo = O.dpm_solver_original_sample(eps_model, noise, K, alg, method, order)
    assert rel_err(x, xo) < tol, rel_err(x, xo)


@pytest.mark.parametrize("method,kw,n", [("rho_rk", dict(), 10), ("rho_rk", dict(), 5), ("rho_rk", dict(rk_method="2heun"), 15),
                                         ("rho_ab", dict(ab_order=3), 15), ("rho_ab", dict(ab_order=2), 10), ("ipndm", dict(), 15)])
def test_deis_matrix_equals_original_sampler(method, kw, n):
    triple = G.deis_triple(n, method, **kw)
    x, noise, eps_model, s = _run(triple, 128, seed=22)
    xo = O.deis_original_sample(eps_model, noise, n, method, **kw)
    assert rel_err(x, xo) < 1e-5, rel_err(x, xo)


def test_original_and_optimised_matrices_share_one_kernel_path(weights_dir):
    """"original sampler vs Natural Inference" on one code path: the generated DPM-Solver++(3M) matrix and the reference's
    optimised step_15_weight_173 on the same 15-node quadratic grid both run as 15 launches of the specialised/generic
    step kernels -- same launch count, dense vs banded history"""
    import os
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200.coeffs import CoeffTriple
    opt = CoeffTriple.from_npz(os.path.join(weights_dir, "step_15_weight_173.npz"))
    orig = G.dpm_solver_triple(15, "dpmsolver++", "multistep", 3)
    assert np.abs(opt.node[:, 0] - orig.node[:, 0]).max() < 1e-6  # the same time grid
    den = lambda x, k: torch.tanh(x) * 0.3
    counts = {}
    for name, t in (("opt", opt), ("orig", orig)):
        s = NaturalInferenceSampler(t, ni.io_score_vp(t.node), 64, (3, 32, 32), device=DEV, seed=1, advance=0)
        n0 = ni.launch_count()
        s.sample(den)
        counts[name] = (ni.launch_count() - n0, s.plan.n_x0_slots, s.plan.total_units(1))
    assert counts["opt"][0] == counts["orig"][0] == 16  # Philox init + 15 steps
    assert counts["opt"][1] == 5 and counts["orig"][1] == 14 and counts["opt"][2] < counts["orig"][2]
