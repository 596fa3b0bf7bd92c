"""Build libni_b200.so in-tree with nvcc for sm_100a (no torch C++ dependency: a plain C-ABI .so).

    python -m naturaldiffusion_b200.build [--force] [-DNAME=VALUE ...] [-o path]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SRCS = [os.path.join(CSRC, n) for n in ("ni_kernels.cu", "ni_step_lean_f32.cu", "ni_step_lean_f16.cu", "ni_step_lean_bf16.cu", "ni_fid.cu")]
SRC = SRCS[0]
HDR = os.path.join(ROOT, "include", "ni_b200.h")
DEPS = SRCS + [HDR, os.path.join(CSRC, "ni_common.cuh"), os.path.join(CSRC, "ni_step_lean.cuh")]
OUT = os.path.join(HERE, "libni_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-diag-suppress", "177",
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libni_b200.so")


def needs_build(out=OUT) -> bool:
    if not os.path.isfile(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(p) > t for p in DEPS + [__file__])


def build(force: bool = False, defines=(), out: str = OUT, verbose: bool = False) -> str:
    if not force and not defines and not needs_build(out):
        return out
    cmd = [find_nvcc(), *NVCC_FLAGS, "--threads", "0", "-I", os.path.join(ROOT, "include"), *defines, "-o", out, *SRCS]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


def build_demo() -> str:
    """examples/c_abi_demo.cu -> build/c_abi_demo (plain C++ client of the C ABI; run by the GPU tests)."""
    src = os.path.join(ROOT, "examples", "c_abi_demo.cu")
    exe = os.path.join(ROOT, "build", "c_abi_demo")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if os.path.isfile(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(p) for p in (src, HDR, OUT)):
        return exe
    cmd = [find_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", HERE, "-lni_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../naturaldiffusion_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on the C-ABI demo:\n" + r.stdout + r.stderr)
    return exe


def build_oracle() -> str:
    """The C part of the CPU oracle (test infrastructure) -- gcc only."""
    d = os.path.join(ROOT, "oracle")
    r = subprocess.run(["make", "-C", d], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return os.path.join(d, "_build", "libni_oracle.so")


if __name__ == "__main__":
    args = sys.argv[1:]
    out = OUT
    if "-o" in args:
        i = args.index("-o")
        out = args[i + 1]
        del args[i:i + 2]
    print(build(force="--force" in args, defines=[a for a in args if a.startswith("-D")], out=out, verbose="-v" in args))
