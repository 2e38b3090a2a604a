#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in "" "SDB_GEMM_TMA_STORE=0" "SDB_PDL=0"; do
env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/tmp_b.json 2> $OUT/tmp_b.err; python - "$cfg" <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_b.json').read().strip().splitlines()[-1])
print(sys.argv[1] or 'default', 'value', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'host enqueue', round(d['e2e']['host_enqueue_ms_per_step'],2))
P
done
