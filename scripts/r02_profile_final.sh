#!/bin/bash
# Re-capture of the lean-kernel evidence with the FINAL library (256-bit instantiations); generic / TMA captures of
# scripts/r02_profile.sh stay valid (those kernels did not change).  Run on the GPU box under gpurun (1 GPU).
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --no-graph --no-cpu-baseline --no-e2e --no-cold --no-per-config"
raw() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_c2_launches.csv $B --config c2 --steps 2 --warmup 1 --block 8 > $O/r02_c2_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_c3_launches.csv $B --config c3 --steps 2 --warmup 1 --block 4 > $O/r02_c3_launches.log 2>&1
ncu --set full --clock-control none -k regex:ni_step_lean -s 40 -c 10 -f -o $O/r02_c2_full $B --config c2 --steps 1 --warmup 1 --block 4 > $O/r02_c2_full.log 2>&1; raw r02_c2_full
ncu --set full --clock-control none --import-source on -k regex:ni_step_lean -s 30 -c 15 -f -o $O/r02_c3_full $B --config c3 --steps 1 --warmup 1 --block 2 > $O/r02_c3_full.log 2>&1; raw r02_c3_full
ncu -i $O/r02_c3_full.ncu-rep --page details --launch-skip 7 --launch-count 1 > $O/r02_c3_step7_details.txt 2>&1
ncu -i $O/r02_c3_full.ncu-rep --page source --csv --launch-skip 7 --launch-count 1 > $O/r02_c3_step7_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:ni_step_lean -s 600 -c 6 -f -o $O/r02_c4_markov_full $B --config c4 --markov 1 --steps 1 --warmup 1 --block 1 > $O/r02_c4_markov_full.log 2>&1; raw r02_c4_markov_full
ncu -i $O/r02_c4_markov_full.ncu-rep --page details --launch-skip 3 --launch-count 1 > $O/r02_c4_markov_details.txt 2>&1
ncu --set full --clock-control none -k regex:ni_step_lean -s 10 -c 2 -f -o $O/r02_pixel_lean python scripts/prof_targets.py pixel --reps 1 > $O/r02_pixel_lean.log 2>&1; raw r02_pixel_lean
ncu --set full --clock-control none -k regex:ni_step_kernel -s 10 -c 2 -f -o $O/r02_pixel_generic python scripts/prof_targets.py pixel --variant 1 --reps 1 > $O/r02_pixel_generic.log 2>&1; raw r02_pixel_generic
find $O -name '*.ncu-rep' -delete
find $O -size +6M -exec gzip -f {} \;
find $O -size +12M -delete
du -sh $O
