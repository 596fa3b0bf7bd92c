"""naturaldiffusion_b200 -- B200 (sm_100a) Natural Inference sampling step.

Only what the hot path needs: the ctypes binding of libni_b200.so (``_lib``), tensor-level wrappers
(``ops``), coefficient matrices + launch plan (``coeffs``), the sampler loop with its device ring
buffer (``sampler``) and the function-level drop-ins for the reference scripts (``dropin``).
Importing the package does not load CUDA; the first compute call does, and fails loudly if the
extension is not built.
"""
from ._lib import NiError, launch_count, lib  # noqa: F401
from .coeffs import CoeffTriple, build_plan, io_eps_cfg, io_score_vp, io_velocity_cfg  # noqa: F401

__all__ = ["NiError", "lib", "launch_count", "CoeffTriple", "build_plan", "io_eps_cfg", "io_score_vp", "io_velocity_cfg"]
