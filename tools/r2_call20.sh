#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for mask in 0xffff 0xfffe 0xfff8 0xffe0 0xff00 0x00ff 0x0007 0x0001; do
SDB_FB_LEVELS=$mask timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/tmp_b.json 2> $OUT/tmp_b.err; python - $mask <<'P'
import json,sys
d=json.loads(open('gpurun_out/tmp_b.json').read().strip().splitlines()[-1])
print('levels',sys.argv[1],'render bwd kernels alone', round(d['profile']['render_bwd_kernel_ms'],2), 'kept', d['profile']['render_samples_kept'])
P
done
