#!/usr/bin/env python
"""Registers / spills per kernel from `nvcc -Xptxas -v` (reads stderr text on stdin or compiles the given .cu files)."""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "-diag-suppress", "177", "-Xptxas", "-v", "-I", os.path.join(ROOT, "include")]

def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"ni::\(anonymous namespace\)::|\(anonymous namespace\)::", "", o) for o in out]

def report(text):
    rows, cur = [], None
    for ln in text.split("\n"):
        m = re.search(r"Compiling entry function '([^']+)'", ln)
        if m:
            cur = [m.group(1), None, 0, 0]
            rows.append(cur)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", ln)
        if m and cur:
            cur[2], cur[3] = int(m.group(1)), int(m.group(2))
        m = re.search(r"Used (\d+) registers", ln)
        if m and cur:
            cur[1] = int(m.group(1))
    names = demangle([r[0] for r in rows])
    for n, r in zip(names, rows):
        n = re.sub(r"\(.*", "", n.replace("void ", ""))
        print(f"{r[1]:>4} regs {r[2]:>5} stack {r[3]:>5} spill  {n}")

if __name__ == "__main__":
    if len(sys.argv) > 1:
        defs = [a for a in sys.argv[1:] if a.startswith("-D")]
        for src in [a for a in sys.argv[1:] if not a.startswith("-D")]:
            r = subprocess.run(["nvcc", *FLAGS, *defs, "-c", src, "-o", "/dev/null"], capture_output=True, text=True)
            report(r.stderr)
    else:
        report(sys.stdin.read())
