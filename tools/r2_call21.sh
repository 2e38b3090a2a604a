#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_render_gpu.py -x -q -s > $OUT/r2p_tests.log 2>&1; grep -E "^table|^w1|^w2|orient only|passed|failed|Error|assert" $OUT/r2p_tests.log | tail -10
for rep in 1 0; do
SDB_FB_REPLICAS=$rep timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r2p_bench_rep$rep.json 2> $OUT/tmp_b.err; python - $rep <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/r2p_bench_rep{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('replicas',sys.argv[1],'C2', round(d['value'],2), round(d['ms_per_step'],2),'render bwd kernels alone', round(d['profile']['render_bwd_kernel_ms'],2), 'kept', d['profile']['render_samples_kept'])
P
done
