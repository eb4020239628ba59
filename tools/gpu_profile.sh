#!/bin/bash
# Profiles for profiles/: launch list of the bench command + ncu --set full of the two hot kernels.
#   tools/gpu_profile.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --at-scale none"
$CMD > gpurun_out/plain_list_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}_c2.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
for wl in c2 c3-shard c4-shard; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --at-scale none --no-graph --only-device"
  $CMD > gpurun_out/plain_full_${wl}_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'score_group|prep_kernel' -s 6 -c 2 \
      -o gpurun_out/prof_${TAG}_$wl -f $CMD > gpurun_out/ncu_full_${wl}_$TAG.log 2>&1
  echo "ncu full $wl rc=$?"
done
