#!/bin/bash
# One GPU session: parity tests, the three bench workloads, then (optionally) ncu captures.
#   tools/gpu_check.sh <tag> [ncu]
set -u
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_$TAG.log
for wl in c2 c3-shard c4-shard; do
  steps=200; [ $wl != c2 ] && steps=20
  python bench.py --workload $wl --steps $steps --warmup 5 --at-scale none > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
  echo "bench $wl rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${wl}_$TAG.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("  value %.3e  ms/step %.4f  e2e %.3e (%.3f ms, staged %.3f ms)  pass roofline %.1f GB/s frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["staged_copies_ms_per_step"], r["achieved"], r["frac"]))
    for k in r.get("kernels", []): print("  ", k["kernel"], "%.4f ms  %.1f GB/s frac %.3f" % (k["ms"], k["achieved"], k["frac"]))
except Exception as e: print("  parse failed", e)
PY
done
if [ "${2:-}" = "ncu" ]; then
  python bench.py --workload c3-shard --steps 2 --warmup 3 --no-cpu-baseline --at-scale none --no-graph > gpurun_out/plain_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'score_group|prep_kernel' -s 12 -c 3 \
      -o gpurun_out/prof_${TAG}_c3 -f python bench.py --workload c3-shard --steps 2 --warmup 3 --no-cpu-baseline --at-scale none --no-graph > gpurun_out/ncu_$TAG.log 2>&1
  echo "ncu c3 rc=$?"
  python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --at-scale none --no-graph > gpurun_out/plain2_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c2.csv \
      python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --at-scale none --no-graph > gpurun_out/ncu2_$TAG.log 2>&1
  echo "ncu launches c2 rc=$?"
fi
