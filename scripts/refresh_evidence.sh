#!/bin/bash
# First GPU call of the next session (one B200, about 5 minutes): re-take every piece of evidence under profiles/ that predates the
# per-launch load flavours, with the library as shipped.   gpurun --timeout 600 -- 'bash scripts/refresh_evidence.sh'
set -u
mkdir -p gpurun_out
( timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/refresh_pytest.txt
timeout 150 python bench.py > gpurun_out/refresh_bench_c2.json 2> gpurun_out/refresh_bench_c2.err
bash scripts/scale_all.sh 1                     # every BASELINE shape, 1 GPU -> gpurun_out/scale_n1.jsonl
bash scripts/gpu_profile.sh c3 --config c3      # launch list + ncu --set full of C3 -> gpurun_out/c3_*
timeout 120 python bench.py --config c2 --eager-comparator --no-cpu-baseline --no-e2e > gpurun_out/refresh_c2_eager.json 2>/dev/null
timeout 120 python bench.py --config c3 --eager-comparator --no-cpu-baseline --no-e2e --steps 300 > gpurun_out/refresh_c3_eager.json 2>/dev/null
timeout 60 python examples/sd3_pipeline.py --small | tee gpurun_out/refresh_sd3_example.txt
