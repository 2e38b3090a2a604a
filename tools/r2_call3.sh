#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_guidance_gpu.py tests/test_nets_gpu.py -x -q -s -k "survives or unet or vae" > $OUT/r2c_tests.log 2>&1
grep -E "rel_l2|passed|failed|Error|error" $OUT/r2c_tests.log | tail -12
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2c_bench_c2.json 2> $OUT/r2c_bench_c2.err; tail -c 1500 $OUT/r2c_bench_c2.json; tail -5 $OUT/r2c_bench_c2.err
timeout 400 python bench.py --workload C4 --steps 6 --warmup 3 > $OUT/r2c_bench_c4.json 2> $OUT/r2c_bench_c4.err; tail -c 3000 $OUT/r2c_bench_c4.json; tail -5 $OUT/r2c_bench_c4.err
timeout 400 python bench.py --workload C5 --steps 6 --warmup 3 > $OUT/r2c_bench_c5.json 2> $OUT/r2c_bench_c5.err; tail -c 3000 $OUT/r2c_bench_c5.json; tail -5 $OUT/r2c_bench_c5.err
