#!/bin/bash
# the other BASELINE configurations through bench.py (one JSON line each), end of round 2
OUT=gpurun_out; mkdir -p $OUT
for wl in C3 C4 C5; do
  K=10; [ $wl = C5 ] && K=4
  timeout 900 python bench.py --workload $wl --steps $K --warmup 3 --no-cpu-baseline > $OUT/r2_bench_$wl.json 2> $OUT/r2_bench_$wl.err
  python - $wl <<'P'
import json,sys
wl=sys.argv[1]
d=json.loads(open(f'gpurun_out/r2_bench_{wl}.json').read().strip().splitlines()[-1])
print(wl, 'steps/s', round(d['value'],3), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],3), 'launches/step', d['gpu_launches_per_step'], 'phases', {k: round(v,1) for k,v in d['profile']['phase_ms'].items()})
for r in [d['roofline']]+d['roofline_other_kernels']: print('   ', r['kernel'][:64], r['bound'], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],3), 'ms', round(r['ms_per_step'],2))
P
done
