"""A numpy-backed stand-in for the handful of jax calls th_deis makes, so that the reference's OWN DEIS code
(/root/reference/deps/th_deis, third party to the path, written against jax which is not installed here) can be executed in
the build container to generate golden coefficient matrices.  TEST INFRASTRUCTURE ONLY, used by tests/golden/make_golden.py.

Differences from real jax, stated: arithmetic is float64 (jax defaults to float32, which is where the 3e-6 noise of the
shipped results/deis matrices comes from); `grad` is a complex-step derivative (exact to rounding for the analytic
log-alpha functions th_deis differentiates); `vmap` is a Python loop; `jit` is the identity."""
from __future__ import annotations

import sys
import types

import numpy as np


class JArr(np.ndarray):
    """ndarray with jax's functional `x.at[idx].set(v)`"""

    @property
    def at(self):
        arr = self

        class _At:
            def __getitem__(self, idx):
                class _Set:
                    def set(self, v):
                        out = np.array(arr, copy=True)
                        out[idx] = v
                        return _wrap(out)
                return _Set()
        return _At()


def _wrap(x):
    if isinstance(x, np.ndarray):
        return x.view(JArr)
    if isinstance(x, (np.floating, np.integer, np.complexfloating)):
        return np.asarray(x).view(JArr)
    return x


def _w(f):
    def g(*a, **k):
        return _wrap(f(*a, **k))
    g.__name__ = getattr(f, "__name__", "fn")
    return g


def make_modules():
    jnp = types.ModuleType("jax.numpy")
    for name in ("sqrt", "arange", "power", "concatenate", "linspace", "prod", "log", "zeros", "sum", "flip", "exp", "where", "ones", "cos",
                 "clip", "arccos", "searchsorted", "stack", "abs"):
        setattr(jnp, name, _w(getattr(np, name)))
    jnp.asarray = lambda x, dtype=None: _wrap(np.asarray(x, dtype=(np.float64 if dtype is float else dtype)))
    jnp.shape, jnp.ndim, jnp.pi = np.shape, np.ndim, np.pi
    jnp.float32, jnp.float64 = np.float64, np.float64

    jax = types.ModuleType("jax")
    jax.numpy = jnp

    def vmap(fn, in_axes=0, out_axes=0):
        def mapped(*args):
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
            n = next(len(a) for a, ax in zip(args, axes) if ax is not None)
            outs = [fn(*[(a[i] if ax is not None else a) for a, ax in zip(args, axes)]) for i in range(n)]
            return _wrap(np.stack([np.asarray(o) for o in outs]))
        return mapped

    def grad(fn):
        def d(x):
            h = 1e-30
            return _wrap(np.imag(np.asarray(fn(np.asarray(x, dtype=np.complex128) + 1j * h))) / h)
        return d

    jax.vmap, jax.grad, jax.jit = vmap, grad, (lambda f: f)
    return jax, jnp


def install():
    """put the stand-ins into sys.modules (the caller removes them again)"""
    jax, jnp = make_modules()
    sys.modules["jax"], sys.modules["jax.numpy"] = jax, jnp
    return jax, jnp
