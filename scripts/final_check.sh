set -u
mkdir -p gpurun_out
( timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/final_pytest_gpu.txt
( timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/final_smoke.txt
timeout 200 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>&1
timeout 300 python scripts/sweep.py --ldg-only --lib naturaldiffusion_b200/libni_b200.so --lib build/alt/lib_st1.so --lib build/alt/lib_ld0.so --lib build/alt/lib_ld2.so --lib build/alt/lib_ld2st1.so > gpurun_out/final_policy_sweep.txt 2>&1
cat gpurun_out/final_pytest_gpu.txt gpurun_out/final_smoke.txt gpurun_out/final_policy_sweep.txt
