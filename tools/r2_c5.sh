#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT


timeout 900 python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/r2v_bench_c5.json 2> $OUT/tmp_c5.err; tail -3 $OUT/tmp_c5.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2v_bench_c5.json').read().strip().splitlines()[-1])
print('C5 step ms', round(d['ms_per_step'],1), 'value', round(d['value'],3), {k:v for k,v in d.get('profile',{}).items() if 'phase' in k})
for r in [d['roofline']]+d['roofline_other_kernels']: print(' ', r['kernel'][:70], r['bound'], round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],3), 'ms', round(r['ms_per_step'],1))
calls=d['profile'].get('abi_calls',{})
for k,v in sorted(calls.items(), key=lambda kv:-kv[1]['ms_per_step'])[:12]: print('   ', k, round(v['ms_per_step'],1))
P
