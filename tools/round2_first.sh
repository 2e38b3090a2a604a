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
