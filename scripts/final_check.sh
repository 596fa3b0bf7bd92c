# Round-end validation on one B200, most important first; every step has its own timeout and the ncu steps are skipped when
# the call is running late (the GPU budget clamps the whole call).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo $(( $(date +%s) - T0 )); }
( timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final_pytest_gpu.txt; echo "pytest done at $(el)s"; cat gpurun_out/final_pytest_gpu.txt
timeout 150 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err; echo "bench c2 done at $(el)s"
: > gpurun_out/final_bench_others.jsonl
for cfg in "c5s --steps 300" "c5 --markov 0 --steps 300" ; do
  timeout 60 python bench.py --config $cfg --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 >> gpurun_out/final_bench_others.jsonl
done
timeout 90 python bench.py --config c3 --steps 300 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/final_bench_others.jsonl; echo "other benches done at $(el)s"
( timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/final_smoke.txt; cat gpurun_out/final_smoke.txt; echo "smoke done at $(el)s"
if [ $(el) -lt 100 ]; then
  timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_c2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-cold > gpurun_out/final_c2_launches.log 2>&1
  echo "launch list done at $(el)s"
fi
if [ $(el) -lt 120 ]; then
  timeout 80 ncu --set full --clock-control none --import-source on -k regex:ni_step_kernel -s 45 -c 10 -f -o gpurun_out/final_c2_full \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-cold > gpurun_out/final_c2_full.log 2>&1
  echo "full capture done at $(el)s"
fi
python - <<'PY'
import json
for f in ("gpurun_out/final_bench_c2.json", "gpurun_out/final_bench_others.jsonl"):
    for ln in open(f):
        try:
            j = json.loads(ln); r = j["roofline"]
            print(j["config"]["workload"][:44], "ms %.4f" % j["ms_per_step"], "GB/s %d" % r["achieved"], "frac %.3f" % r["frac"], "e2e", j["e2e"] and "%.4g" % j["e2e"]["value"])
        except Exception as e:
            print("bad line", ln[:80])
PY
