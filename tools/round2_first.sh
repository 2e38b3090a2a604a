#!/bin/bash
# First GPU call of the next round: the device tests that were written after the round-1 GPU budget was spent, then the
# whole GPU suite, then the evidence capture.  bash tools/round2_first.sh <tag>
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
SDB_UNVERIFIED_TESTS=1 timeout 600 python -m pytest tests/test_zz_multiprompt_eval_gpu.py -x -q > $OUT/${TAG}_unverified.log 2>&1
tail -15 $OUT/${TAG}_unverified.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1
tail -5 $OUT/${TAG}_gpu_tests.log
bash tools/capture_evidence.sh $TAG
# CTA-pair tile probe (bounded waits; a protocol mistake prints TIMEOUT, it cannot hang)
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -I scaledreamer_b200/csrc tools/gemm2cta_probe.cu \
  -o $OUT/gemm2cta_probe && timeout 60 $OUT/gemm2cta_probe > $OUT/${TAG}_gemm2cta_probe.log 2>&1
cat $OUT/${TAG}_gemm2cta_probe.log
