#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_amortized_gpu.py tests/test_guidance_gpu.py -x -q -s -k "chunked or guidance_matches or parity" > $OUT/r2d_tests.log 2>&1
grep -E "guidance parity|passed|failed|Error|error|assert" $OUT/r2d_tests.log | tail -12
timeout 200 python tools/gemm_timeline2.py > $OUT/r2d_timeline2.log 2>&1; cat $OUT/r2d_timeline2.log
timeout 500 python bench.py --workload C5 --steps 4 --warmup 3 > $OUT/r2d_bench_c5.json 2> $OUT/r2d_bench_c5.err; tail -c 2500 $OUT/r2d_bench_c5.json; tail -5 $OUT/r2d_bench_c5.err
