#!/bin/bash
# Round-2 evidence for profiles/: launch list of the bench command, ncu --set full of the hot kernels (the one-launch kernel
# on C2, pack + score on the C3 / C4 shards), ncu of the kernels beside the scoring pass, and their event-timed rooflines.
#   tools/gpu_profile_r02.sh      (every ncu command runs only after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --at-scale none --strong none"
$CMD > gpurun_out/plain_list_r02.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_c2.csv $CMD > gpurun_out/ncu_list_r02.log 2>&1
echo "launch list rc=$?"
for wl in c2 c3-shard c4-shard; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --at-scale none --strong none --no-graph --only-device"
  $CMD > gpurun_out/plain_full_${wl}_r02.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'fused_score|score_group|prep_kernel' -s 3 -c 2 \
      -o gpurun_out/prof_r02_$wl -f $CMD > gpurun_out/ncu_full_${wl}_r02.log 2>&1
  echo "ncu full $wl rc=$?"
done
python tools/exp_aux_kernels.py gpurun_out/r02_aux_kernels.json > gpurun_out/aux_plain.log 2>&1 &&
AUX_ONCE=1 ncu --set full --clock-control none -k regex:'emit_|ingest_|match_kernel|snp_|scan_' -c 24 -o gpurun_out/prof_r02_aux -f \
    python tools/exp_aux_kernels.py > gpurun_out/ncu_aux.log 2>&1
echo "aux rc=$?"
python tools/refresh_profiles_r02.py > gpurun_out/refresh_r02.log 2>&1
echo "refresh rc=$?"
du -sh gpurun_out
