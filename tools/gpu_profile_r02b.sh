#!/bin/bash
# Round-2 evidence, final state: the bench command's headline is a STREAM of independent one-launch passes
# (fused_score_pair_kernel).  Launch list of that command + ncu --set full of the kernel (under ncu the launches are
# serialised, so a launch shows the kernel ALONE on the machine; the bench's per-pass time is two CTAs per SM overlapped).
#   tools/gpu_profile_r02b.sh     (every ncu command runs only after the same command exited 0 without ncu)
mkdir -p gpurun_out/profiles_r02b
CMD="python bench.py --steps 120 --warmup 5 --no-cpu-baseline --at-scale none --strong none"
$CMD > gpurun_out/plain_list_r02b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/profiles_r02b/r02_launches_c2_stream.csv $CMD > gpurun_out/ncu_list_r02b.log 2>&1
echo "launch list rc=$?"
CMD="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --at-scale none --strong none --only-device"
$CMD > gpurun_out/plain_full_r02b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fused_score_pair' -s 20 -c 2 \
    -o gpurun_out/prof_r02b_c2 -f $CMD > gpurun_out/ncu_full_r02b.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/prof_r02b_c2.ncu-rep > gpurun_out/profiles_r02b/r02_ncu_full_c2_stream.txt
python tools/ncu_lines.py gpurun_out/prof_r02b_c2.ncu-rep fused_score_pair 1.0 > gpurun_out/profiles_r02b/r02_ncu_lines_fused_pair_c2.txt
rm -f gpurun_out/prof_r02b_c2.ncu-rep
tail -3 gpurun_out/plain_list_r02b.log | cut -c1-600
