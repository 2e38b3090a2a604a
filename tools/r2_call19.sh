#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for tc in 1 0; do for ns in 0 1; do
if [ $ns = 1 ]; then export SDB_FB_NOSCATTER=1; else unset SDB_FB_NOSCATTER; fi
SDB_FB_TC=$tc timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/tmp_b.json 2> $OUT/tmp_b.err; python - $tc $ns <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_b.json').read().strip().splitlines()[-1])
print('TC',sys.argv[1],'noscatter',sys.argv[2],'render bwd kernels alone', round(d['profile']['render_bwd_kernel_ms'],2), 'fwd', round(d['profile']['render_fwd_kernel_ms'],2), 'kept', d['profile']['render_samples_kept'])
P
done; done
