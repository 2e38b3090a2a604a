#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_nn_kernels_gpu.py -x -q > $OUT/r2h_tests.log 2>&1
tail -5 $OUT/r2h_tests.log
timeout 300 python tools/gemm_shapes.py streamk > $OUT/r2h_gemm_streamk.log 2>&1; cat $OUT/r2h_gemm_streamk.log
SDB_GEMM_STREAMK=0 timeout 300 python tools/gemm_shapes.py nostreamk > $OUT/r2h_gemm_nostreamk.log 2>&1; tail -1 $OUT/r2h_gemm_nostreamk.log
timeout 300 python -m pytest tests/test_nets_gpu.py -x -q -s 2>&1 | grep -E "rel_l2|passed|failed"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2h_bench_c2.json 2> $OUT/r2h_bench_c2.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2h_bench_c2.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['ms_per_step'], 'gemm ms', d['profile']['gemm_ms_per_step'], 'frac', [ (r['kernel'][:20], round(r['frac'],3), round(r['ms_per_step'],2)) for r in [d['roofline']]+d['roofline_other_kernels']])
P
