#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for i in 1 2 3 4; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/tmp_b.json 2> $OUT/tmp_b.err; python - <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_b.json').read().strip().splitlines()[-1])
print('value', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'host enqueue', round(d['e2e']['host_enqueue_ms_per_step'],2), d['clocks'])
P
done
