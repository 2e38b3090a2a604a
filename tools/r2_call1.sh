#!/bin/bash
# round 2, first device call: the multi-prompt eval test that was skipped, the CTA-pair probe, then the GPU suite
OUT=gpurun_out; mkdir -p $OUT
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -I scaledreamer_b200/csrc tools/gemm2cta_probe.cu \
  -o $OUT/gemm2cta_probe && timeout 60 $OUT/gemm2cta_probe > $OUT/r2a_gemm2cta_probe.log 2>&1
echo "probe rc=$?"; cat $OUT/r2a_gemm2cta_probe.log | tail -30
SDB_UNVERIFIED_TESTS=1 timeout 600 python -m pytest tests/test_zz_multiprompt_eval_gpu.py -x -q > $OUT/r2a_unverified.log 2>&1
tail -25 $OUT/r2a_unverified.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2a_gpu_tests.log 2>&1
tail -5 $OUT/r2a_gpu_tests.log
