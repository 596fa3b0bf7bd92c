#!/usr/bin/env python
"""SASS opcode histogram per kernel of libni_b200.so (CPU side: cuobjdump).  Evidence for which instructions the
kernels actually issue: LDG.E.128 / LDG.E.NA.128 / STG.E.128 (vector global access), UBLKCP + SYNCS (TMA bulk copies and
mbarriers), ACQBULK / PREEXIT (programmatic dependent launch), DMMA (fp64 tensor cores), MUFU (SFU), IMAD.WIDE (Philox).

    python scripts/sass_histogram.py [regex ...] > profiles/r02_sass_histograms.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "naturaldiffusion_b200", "libni_b200.so")
DEFAULT = [r"ni_step_lean_kernel<float, float, 4, 0, 1, 0, false, 4, 32>", r"ni_step_lean_kernel<float, float, 4, 0, 1, 0, false, 4, 16>",
           r"ni_step_lean_kernel<float, float, 0, 1, 2, 1, false, 1, 32>", r"ni_step_lean_kernel<float, float, -1, 0, 1, 0, false, 32, 32>",
           r"ni_step_lean_kernel<__half, __half, 1, 0, 2, 1, false, 1, 16>", r"ni_step_lean_kernel<__half, __half, -1, 0, 2, 0, false, 32, 32>",
           r"ni_step_lean_kernel<float, float, 3, 0, 1, 0, true", r"ni_step_kernel<float, float, 4, 32, true>",
           r"ni_step_tma_kernel<float, 2048>", r"ni_step_tma_kernel<float, 4096>", r"ni_wsum_kernel<float, float, 4, 32, 0>",
           r"ni_wsum_kernel<double, float, 2, 32, 0>", r"ni_normal_kernel<float, 4>", r"ni_fid_syrk_kernel", r"ni_pixel_kernel<float>"]


def main():
    pats = sys.argv[1:] or DEFAULT
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    names = [f.split("\n")[0].strip() for f in funcs]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    dem = [re.sub(r"ni::\(anonymous namespace\)::|\(anonymous namespace\)::|void ", "", d) for d in dem]
    print(f"# SASS opcode histograms, {os.path.relpath(SO, ROOT)} (sm_100a), {len(funcs)} kernels in the library\n")
    for pat in pats:
        for f, d in zip(funcs, dem):
            if pat not in d:
                continue
            ops = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", f)
            ops = [o for o in ops if o not in ("NOP",)]
            ops = [re.sub(r"(LDG|STG)\.E((?:\.NA)?)\.ENL2\.256", r"\1.E\2.ENL2.256", o) for o in ops]
            base = collections.Counter(re.sub(r"^(LDG\.E(?:\.NA)?(?:\.ENL2)?(?:\.\d+)?|STG\.E(?:\.ENL2)?(?:\.\d+)?|LDS(?:\.\d+)?|MUFU\.\w+|IMAD\.WIDE(?:\.U32)?|DMMA\.\d+|UBLKCP\.\w+\.\w+|SYNCS\.\w+(?:\.\w+)?|RED\.\w+|REDG\.\w+)?.*", lambda m: m.group(1) or m.group(0).split(".")[0], o) for o in ops)
            print(f"## {re.sub(r'[(].*', '', d)}\n   static instructions: {len(ops)}")
            print("   " + ", ".join(f"{k} {v}" for k, v in base.most_common(28)))
            full = collections.Counter(ops)
            key = {k: v for k, v in full.items() if re.match(r"(LDG|STG|LDS|STS|LDL|STL|UBLKCP|SYNCS|ACQBULK|PREEXIT|DMMA|LDGSTS|RED|ATOM|MUFU|UTMA|LDTM|UTC)", k)}
            print("   memory / async / special: " + ", ".join(f"{k} {v}" for k, v in sorted(key.items())) + "\n")
            break
        else:
            print(f"## (no kernel matches {pat!r})\n")


if __name__ == "__main__":
    main()
