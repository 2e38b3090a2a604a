#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_amortized_gpu.py tests/test_packed_gpu.py tests/test_zz_multiprompt_eval_gpu.py tests/test_fullsize_parity_gpu.py -x -q 2>&1 | tail -4
SDB_MLP3_TC=0 timeout 600 python -m pytest tests/test_amortized_gpu.py -x -q -k tiny_mlp 2>&1 | tail -1
timeout 900 python bench.py --workload C5 --steps 4 --warmup 3 --no-cpu-baseline > $OUT/r2_bench_C5.json 2> $OUT/r2_bench_C5.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_C5.json").read().strip().splitlines()[-1])
c=d["profile"]["abi_calls"]
print("C5 ms", round(d["ms_per_step"],1), "steps/s", round(d["value"],3), {k: round(c[k]["ms_per_step"],1) for k in ("sdb_triplane_sample_backward","sdb_triplane_sample_forward","sdb_mlp3_backward","sdb_mlp3_forward")})
P
