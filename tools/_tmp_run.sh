python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -3
for wl in c3-shard c4-shard; do python tools/exp_missing_rates.py $wl 2>&1 | grep -v build; done
for wl in c3-shard c4-shard; do
  python bench.py --workload $wl --at-scale none --strong none --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('$wl pass_ms', round(d['ms_per_step'],5), [(k['kernel'], round(k['ms'],4)) for k in r['kernels']], d['parity']['bit_exact_vs_oracle'])
"
done
