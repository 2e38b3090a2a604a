#!/bin/bash
# whole GPU suite + smoke + a 20-step bench line
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2s}
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; tail -4 $OUT/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_c2.json 2> $OUT/${TAG}_bench.err; python - $TAG <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/{sys.argv[1]}_bench_c2.json').read().strip().splitlines()[-1])
print('C2', round(d['value'],2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'gemm ms', round(d['profile']['gemm_ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'launches/step', d['gpu_launches_per_step'], 'cpu', d['cpu_baseline'])
P
