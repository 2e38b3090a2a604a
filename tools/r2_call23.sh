#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_nn_kernels_gpu.py tests/test_nets_gpu.py -x -q -s 2>&1 | grep -E "rel_l2|passed|failed" | tail -5
timeout 300 python tools/gemm_shapes.py sbias > $OUT/r2r_gemm_sbias.log 2>&1; cat $OUT/r2r_gemm_sbias.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r2r_bench_c2.json 2> $OUT/tmp_b.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2r_bench_c2.json').read().strip().splitlines()[-1])
print('C2', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'gemm ms', round(d['profile']['gemm_ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))
P
