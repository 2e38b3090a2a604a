#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_nn_kernels_gpu.py tests/test_nets_gpu.py tests/test_guidance_gpu.py -x -q -s > $OUT/r2m_tests.log 2>&1; grep -E "rel_l2|guidance parity|passed|failed" $OUT/r2m_tests.log | tail -6
SDB_GEMM_CSV=$OUT/r2m_gemm_per_launch.csv timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2m_bench_c2.json 2> $OUT/r2m_bench_c2.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2m_bench_c2.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['ms_per_step'], 'gemm ms', d['profile']['gemm_ms_per_step'], 'frac', [ (r['kernel'][:20], round(r['frac'],3), round(r['ms_per_step'],2)) for r in [d['roofline']]+d['roofline_other_kernels']])
P
