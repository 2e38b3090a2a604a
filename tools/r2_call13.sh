#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_nn_kernels_gpu.py tests/test_nets_gpu.py -x -q -s > $OUT/r2j_tests.log 2>&1; grep -E "rel_l2|passed|failed" $OUT/r2j_tests.log | tail -5
timeout 300 python tools/gemm_shapes.py producer > $OUT/r2j_gemm_producer.log 2>&1; cat $OUT/r2j_gemm_producer.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2j_bench_c2.json 2> $OUT/r2j_bench_c2.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2j_bench_c2.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['ms_per_step'], 'gemm ms', d['profile']['gemm_ms_per_step'], 'frac', [ (r['kernel'][:20], round(r['frac'],3), round(r['ms_per_step'],2)) for r in [d['roofline']]+d['roofline_other_kernels']])
print({k:round(v['ms_per_step'],2) for k,v in d['profile']['abi_calls'].items() if v['ms_per_step']>0.1})
P
