#!/bin/bash
# Round-2 evidence, run on the GPU box under gpurun (1 GPU):  scripts/r02_profile.sh
# launch lists (shares) + `ncu --set full` of the dominant kernels, exported to CSV ON THE BOX; no .ncu-rep comes back
# (gpurun_out/ is capped at 64 MiB).
set -u
O=gpurun_out; mkdir -p $O
B="python bench.py --no-graph --no-cpu-baseline --no-e2e --no-cold --no-per-config"
raw() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null; }
# launch lists of the timed region (2 steps of 8 trajectories after warm-up)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_c2_launches.csv $B --config c2 --steps 2 --warmup 1 --block 8 > $O/r02_c2_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_c3_launches.csv $B --config c3 --steps 2 --warmup 1 --block 4 > $O/r02_c3_launches.log 2>&1
# full sets: one whole trajectory of each
ncu --set full --clock-control none -k regex:ni_step_lean -s 40 -c 10 -f -o $O/r02_c2_full $B --config c2 --steps 1 --warmup 1 --block 4 > $O/r02_c2_full.log 2>&1; raw r02_c2_full
ncu --set full --clock-control none --import-source on -k regex:ni_step_lean -s 30 -c 15 -f -o $O/r02_c3_full $B --config c3 --steps 1 --warmup 1 --block 2 > $O/r02_c3_full.log 2>&1; raw r02_c3_full
ncu -i $O/r02_c3_full.ncu-rep --page details --launch-skip 7 --launch-count 1 > $O/r02_c3_step7_details.txt 2>&1
ncu -i $O/r02_c3_full.ncu-rep --page source --csv --launch-skip 7 --launch-count 1 > $O/r02_c3_step7_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:ni_step_lean -s 600 -c 6 -f -o $O/r02_c4_markov_full $B --config c4 --markov 1 --steps 1 --warmup 1 --block 1 > $O/r02_c4_markov_full.log 2>&1; raw r02_c4_markov_full
ncu -i $O/r02_c4_markov_full.ncu-rep --page details --launch-skip 3 --launch-count 1 > $O/r02_c4_markov_details.txt 2>&1
ncu --set full --clock-control none -k regex:ni_step_kernel -s 600 -c 6 -f -o $O/r02_c4_markov_generic_full $B --variant 1 --config c4 --markov 1 --steps 1 --warmup 1 --block 1 > $O/r02_c4_markov_generic_full.log 2>&1; raw r02_c4_markov_generic_full
# generic kernel and TMA kernel on the same C3 steps (side by side with the lean one)
ncu --set full --clock-control none -k regex:ni_step_kernel -s 36 -c 6 -f -o $O/r02_c3_generic_full $B --variant 1 --config c3 --steps 1 --warmup 1 --block 2 > $O/r02_c3_generic_full.log 2>&1; raw r02_c3_generic_full
ncu --set full --clock-control none -k regex:ni_step_tma -s 36 -c 6 -f -o $O/r02_c3_tma_full $B --variant 2 ${TMA_OPTS:-} --config c3 --steps 1 --warmup 1 --block 2 > $O/r02_c3_tma_full.log 2>&1; raw r02_c3_tma_full
ncu -i $O/r02_c3_tma_full.ncu-rep --page details --launch-skip 1 --launch-count 1 > $O/r02_c3_tma_step7_details.txt 2>&1
# the pixel stage (lean vs generic), the stand-alone weighted sum, the FID rank-k update
ncu --set full --clock-control none -k regex:ni_step_lean -s 24 -c 2 -f -o $O/r02_pixel_lean python scripts/prof_targets.py pixel --reps 1 > $O/r02_pixel_lean.log 2>&1; raw r02_pixel_lean
ncu --set full --clock-control none -k regex:ni_step_kernel -s 24 -c 2 -f -o $O/r02_pixel_generic python scripts/prof_targets.py pixel --variant 1 --reps 1 > $O/r02_pixel_generic.log 2>&1; raw r02_pixel_generic
ncu --set full --clock-control none -k regex:ni_wsum -c 6 -f -o $O/r02_wsum python scripts/prof_targets.py wsum --reps 1 > $O/r02_wsum.log 2>&1; raw r02_wsum
ncu --set full --clock-control none -k regex:ni_fid_syrk -c 2 -f -o $O/r02_fid python scripts/prof_targets.py fid --reps 1 > $O/r02_fid.log 2>&1; raw r02_fid
ncu -i $O/r02_fid.ncu-rep --page details --launch-skip 1 --launch-count 1 > $O/r02_fid_details.txt 2>&1
# bare timings of the same targets (no profiler)
for t in pixel wsum fid normal; do python scripts/prof_targets.py $t > $O/r02_target_$t.json 2>$O/r02_target_$t.err; done
python scripts/prof_targets.py pixel --variant 1 > $O/r02_target_pixel_generic.json 2>/dev/null
find $O -name '*.ncu-rep' -delete
find $O -size +6M -exec gzip -f {} \;
find $O -size +12M -delete
du -sh $O
ls -la $O | head -70
