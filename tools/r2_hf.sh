#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in "SDB_HF_TC=1" "SDB_HF_TC=0" "SDB_HF_TC=1" "SDB_HF_TC=0"; do
env $cfg timeout 600 python bench.py --workload C4 --steps 10 --warmup 4 --no-cpu-baseline > $OUT/tmp_c4.json 2> $OUT/tmp_c4.err; python - "$cfg" <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_c4.json').read().strip().splitlines()[-1])
c=d['profile']['abi_calls']
print(sys.argv[1], 'C4 step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'phases', {k: round(v,1) for k,v in d['profile']['phase_ms'].items()}, 'hf bwd', round(c['sdb_hyper_field_backward']['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'tl', round(d['timeline']['compute_ms_mean_over_ranks'],1))
P
done
