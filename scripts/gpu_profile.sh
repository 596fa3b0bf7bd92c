#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the fused step kernel.
# Usage: scripts/gpu_profile.sh <tag> [bench args...]      outputs -> gpurun_out/<tag>_*
set -u
tag=$1; shift
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-cold "$@" > gpurun_out/${tag}_launches.log 2>&1
# the top kernel, full set, source-level
ncu --set full --clock-control none --import-source on -k regex:ni_step_kernel -s 45 -c 15 -f -o gpurun_out/${tag}_full \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-cold "$@" > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out/
