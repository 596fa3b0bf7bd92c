#!/bin/bash
# Weak-scaling table over every BASELINE shape at N GPUs of one box:  scripts/scale_all.sh N  -> gpurun_out/scale_nN.jsonl
N=$1
out=gpurun_out/scale_n${N}.jsonl
mkdir -p gpurun_out; : > $out
port=29600
for cfg in "c2" "c3 --steps 300" "c4 --steps 30" "c4 --markov 0 --steps 4" "c5 --steps 200" "c5s --steps 200" "c5 --batch 256 --markov 0 --steps 50"; do
  port=$((port+1))
  if [ "$N" = "1" ]; then
    python bench.py --config $cfg --no-cpu-baseline 2>/dev/null | tail -1 >> $out
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config $cfg --no-cpu-baseline 2>/dev/null | tail -1 >> $out
  fi
done
python - <<PY
import json
for ln in open("$out"):
    try:
        j=json.loads(ln); r=j["roofline"]
        print(j["n_gpus"], j["config"]["workload"][:48], "markov" if j["config"]["markov_fast_path"] else "dense ", "samples/s %.4g" % j["value"], "GB/s/GPU %d" % r["achieved"], "frac %.3f" % r["frac"], "e2e %.4g" % j["e2e"]["value"])
    except Exception as e:
        print("bad line", ln[:100])
PY
