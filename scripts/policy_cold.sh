# Next measurement to take (needs a B200): the load-flavour choice under COLD L2 (every launch alone after an L2 flush), on every shape.
# Back-to-back numbers are in profiles/r01_policy_sweep.txt; C2 flips sign between the two regimes.
set -u
mkdir -p gpurun_out; out=gpurun_out/policy_cold.txt; : > $out
for cfg in "c2 --steps 200" "c3 --steps 50" "c4 --markov 0 --steps 2" "c4 --steps 10" "c5 --steps 50" "c5 --markov 0 --steps 50" "c5s --steps 50" "c5 --batch 256 --markov 0 --steps 10"; do
  for pol in 1 2 0; do
    timeout 120 python bench.py --config $cfg --cold --no-e2e --no-cpu-baseline --opt load_policy=$pol 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; c=r['cold_l2']
print('$cfg load_policy=$pol warm ms %.4f GB/s %d | cold ms %.4f GB/s %d' % (j['ms_per_step'], r['achieved'], c['ms_per_trajectory'], c['achieved']))" | tee -a $out
  done
done
