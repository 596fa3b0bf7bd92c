"""Import the reference's OWN functions for the NI path (this container only).

TEST INFRASTRUCTURE ONLY (see oracle/ni_oracle.py).  /root/reference does not exist on
the GPU box; everything here is used solely by ``tests/golden/make_golden.py`` (to
produce committed fixtures) and by CPU tests that skip when the reference is absent.

The reference scripts import third-party modules that are not installed offline
(diffusers, timm, ml_collections, jax, tensorflow, pytorch_fid, matplotlib ...).
Those are stubbed in ``sys.modules`` *only while importing*; the functions we take
are pure torch/numpy.  For ``src/CIFAR10NaturalInference.py`` the module-level import
chain is too wide to stub honestly, so the two functions on the path (``data_fn``,
``weighted_sum``; lines 218-238) are exec'd from their exact source lines instead.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("NI_REFERENCE_ROOT", "/root/reference")


def present() -> bool:
    """the reference tree exists (reading it as TEXT is always allowed: signatures(), the shipped matrices)"""
    return os.path.isfile(os.path.join(REF_ROOT, "src", "ValidateNaturalInference.py"))


def available() -> bool:
    """EXECUTING reference code is opt-in: the tree is untrusted public content, and importing it leaves stubbed
    ``sys.modules`` entries and ``sys.path`` edits behind while it runs.  Only ``tests/golden/make_golden.py`` (which
    sets NI_EXEC_REFERENCE=1 itself) and a developer who exports the flag get the exec-based loaders; default test
    runs rely on the committed goldens."""
    return present() and os.environ.get("NI_EXEC_REFERENCE", "") == "1"


def signatures(relpath: str) -> dict:
    """{function name: [parameter names]} of the top-level functions of a reference source file, by PARSING it
    (ast) -- nothing is executed."""
    import ast
    with open(os.path.join(REF_ROOT, relpath)) as f:
        tree = ast.parse(f.read())
    return {n.name: [a.arg for a in n.args.args] for n in tree.body if isinstance(n, ast.FunctionDef)}


class _Stub(types.ModuleType):
    def __getattr__(self, name):  # any attribute resolves to None-like placeholder
        if name.startswith("__"):
            raise AttributeError(name)
        return None


def _with_stubs(names, fn):
    saved = {n: sys.modules.get(n) for n in names}
    saved_path = list(sys.path)
    try:
        for n in names:
            sys.modules[n] = _Stub(n)
        return fn()
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
        sys.path[:] = saved_path


def _load(path, modname):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def validate_module():
    """src/ValidateNaturalInference.py with diffusers / DiT model imports stubbed."""
    def go():
        return _load(os.path.join(REF_ROOT, "src", "ValidateNaturalInference.py"), "_ref_validate")
    return _with_stubs(["diffusers", "diffusers.models", "models", "torchvision", "torchvision.utils"], go)


def sd3_module():
    """src/SD3NaturalInference.py with diffusers / cv2 / PIL stubbed."""
    def go():
        return _load(os.path.join(REF_ROOT, "src", "SD3NaturalInference.py"), "_ref_sd3")
    return _with_stubs(["diffusers", "cv2", "PIL"], go)


def cifar_functions():
    """``data_fn`` and ``weighted_sum`` exec'd from src/CIFAR10NaturalInference.py:218-238."""
    import numpy as np
    import torch
    path = os.path.join(REF_ROOT, "src", "CIFAR10NaturalInference.py")
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def data_fn(")) - 1  # decorator line
    end = next(i for i, l in enumerate(lines) if l.startswith("def natural_inference_tx("))
    while not lines[end - 1].startswith("@"):
        end -= 1
    src = "\n".join(lines[start:end - 1])
    ns = {"torch": torch, "np": np}
    exec(compile(src, path, "exec"), ns)
    return ns["data_fn"], ns["weighted_sum"]


def analyze_ddpm_ddim_module():
    """src/AnalyzeDDPMDDIM.py (closed-form + sympy generators) with plotting stubbed."""
    def go():
        sys.path.insert(0, os.path.join(REF_ROOT, "src"))
        sys.modules.pop("Utils", None)
        return _load(os.path.join(REF_ROOT, "src", "AnalyzeDDPMDDIM.py"), "_ref_analyze_ddpmddim")
    return _with_stubs(["matplotlib", "matplotlib.pyplot", "scienceplots"], go)


def score_sde_vp():
    """(VPSDE class, get_score_fn) from deps/score_sde_pytorch with ml_collections/op stubbed.
    Only sde_lib + models/utils are needed for the VP score wrapper (models/utils.py:129-160)."""
    def go():
        root = os.path.join(REF_ROOT, "deps", "score_sde_pytorch")
        sys.path.insert(0, root)
        for m in ("sde_lib", "models", "models.utils"):
            sys.modules.pop(m, None)
        sde_lib = _load(os.path.join(root, "sde_lib.py"), "sde_lib")
        sys.modules["sde_lib"] = sde_lib
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(root, "models")]
        sys.modules["models"] = pkg
        mutils = _load(os.path.join(root, "models", "utils.py"), "models.utils")
        sys.modules.pop("models", None)
        sys.modules.pop("sde_lib", None)
        return sde_lib.VPSDE, mutils.get_score_fn
    return _with_stubs(["ml_collections", "op"], go)


def dpm_solver_module():
    """deps/dpm_solver_pytorch.py (pure torch): NoiseScheduleVP, DPM_Solver -- the original samplers behind
    results/FID/dpmsolver*_*.csv."""
    return _load(os.path.join(REF_ROOT, "deps", "dpm_solver_pytorch.py"), "_ref_dpm_solver")


def th_deis_module():
    """deps/th_deis (the DEIS samplers behind results/FID/deis_*step.csv; jax code) imported with oracle/jax_numpy_shim.py standing
    in for jax: numpy float64 arithmetic, complex-step `grad`, loop `vmap`."""
    from . import jax_numpy_shim
    saved = {n: sys.modules.get(n) for n in ("jax", "jax.numpy")}
    saved_path = list(sys.path)
    try:
        jax_numpy_shim.install()
        sys.path.insert(0, os.path.join(REF_ROOT, "deps"))
        for m in [m for m in sys.modules if m == "th_deis" or m.startswith("th_deis.")]:
            sys.modules.pop(m)
        import importlib
        return importlib.import_module("th_deis")
    finally:
        sys.path[:] = saved_path
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
