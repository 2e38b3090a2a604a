#!/bin/bash
# repeated short bench runs: step time spread from run to run
OUT=gpurun_out; mkdir -p $OUT
for i in 1 2 3 4; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/tmp_b.json 2> $OUT/tmp_b.err; python - $i <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_b.json').read().strip().splitlines()[-1])
print('run', sys.argv[1], 'step', round(d['ms_per_step'],2), 'e2e ms', round(1000/d['e2e']['value'],2), 'host max', round(d['e2e']['host_enqueue_ms_max'],1), 'clocks', d['clocks'])
P
done
for i in 1 2; do
timeout 600 python bench.py --workload C4 --steps 10 --warmup 4 --no-cpu-baseline > $OUT/tmp_c4.json 2> $OUT/tmp_c4.err; python - <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_c4.json').read().strip().splitlines()[-1])
print('C4 step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'host max', round(d['e2e']['host_enqueue_ms_max'],1))
P
done
