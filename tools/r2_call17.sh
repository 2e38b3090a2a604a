#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_nn_kernels_gpu.py tests/test_nets_gpu.py tests/test_guidance_gpu.py -x -q -s > $OUT/r2n_tests.log 2>&1; grep -E "rel_l2|guidance parity|passed|failed" $OUT/r2n_tests.log | tail -6
for pdl in 1 0; do
SDB_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r2n_bench_c2_pdl$pdl.json 2> $OUT/r2n_bench_c2_pdl$pdl.err; python - $pdl <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/r2n_bench_c2_pdl{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('PDL',sys.argv[1],'C2', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'gemm ms', round(d['profile']['gemm_ms_per_step'],2), {k:round(v,2) for k,v in d['profile']['phase_ms'].items()})
print('   ', {k:round(v['ms_per_step'],2) for k,v in d['profile']['abi_calls'].items() if v['ms_per_step']>0.1})
P
done
