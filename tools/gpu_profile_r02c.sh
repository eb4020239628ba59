#!/bin/bash
# Round-2 evidence, final state of the launch pipeline on data with missing calls: ncu --set full of prep_kernel and
# score_group_kernel on the C3 shard with 5e-5 and 5e-3 of the target calls set to -1 (the signed body, the
# warp-cooperative redo), summaries + per-line profiles into gpurun_out/profiles_r02e/.
#   tools/gpu_profile_r02c.sh     (every ncu command runs only after the same command exited 0 without ncu)
mkdir -p gpurun_out/profiles_r02e
for rate in 0.00005 0.005; do
CMD="python tools/exp_missing_once.py c3-shard $rate"
$CMD > gpurun_out/plain_missing.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'score_group|prep_kernel' -s 6 -c 2 \
    -o gpurun_out/prof_missing_$rate -f $CMD > gpurun_out/ncu_missing.log 2>&1
echo "ncu missing $rate rc=$?"
python tools/ncu_summary.py gpurun_out/prof_missing_$rate.ncu-rep > gpurun_out/profiles_r02e/r02_ncu_full_c3-shard_missing_$rate.txt
python tools/ncu_lines.py gpurun_out/prof_missing_$rate.ncu-rep score_group 0.7 > gpurun_out/profiles_r02e/r02_ncu_lines_score_group_c3-shard_missing_$rate.txt
python tools/ncu_lines.py gpurun_out/prof_missing_$rate.ncu-rep prep_kernel 1.0 > gpurun_out/profiles_r02e/r02_ncu_lines_prep_c3-shard_missing_$rate.txt
rm -f gpurun_out/prof_missing_$rate.ncu-rep
done
