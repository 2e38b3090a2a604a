#!/bin/bash
# repeated short bench runs: step time spread from run to run
OUT=gpurun_out; mkdir -p $OUT
for i in 1 2 3 4; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/tmp_b.json 2> $OUT/tmp_b.err; python - $i <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_b.json').read().strip().splitlines()[-1])
print('run', sys.argv[1], 'step', round(d['ms_per_step'],2), 'e2e ms', round(1000/d['e2e']['value'],2), 'render bwd alone', round(d['profile']['render_bwd_kernel_ms'],2), 'phases', {k: round(v,2) for k,v in d['profile']['phase_ms'].items()}, 'clocks', d['clocks']['sm_mhz'])
P
done
