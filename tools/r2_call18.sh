#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_guidance_gpu.py -x -q -s > $OUT/r2o_tests.log 2>&1; grep -E "^table|^w1|^w2|orient only|passed|failed|Error|assert" $OUT/r2o_tests.log | tail -16
SDB_FB_TC=0 timeout 600 python -m pytest tests/test_render_gpu.py -x -q -s -k "orientation" 2>&1 | grep -E "^table|^w1|^w2|passed|failed" | head -8
for tc in 1 0; do
SDB_FB_TC=$tc timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r2o_bench_c2_tc$tc.json 2> $OUT/r2o_bench_c2_tc$tc.err; python - $tc <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/r2o_bench_c2_tc{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('TC',sys.argv[1],'C2', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'render bwd kernels alone', round(d['profile']['render_bwd_kernel_ms'],2), 'fwd', round(d['profile']['render_fwd_kernel_ms'],2), 'kept', d['profile']['render_samples_kept'])
P
done
