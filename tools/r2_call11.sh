#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_nn_kernels_gpu.py -x -q -k "conv" > $OUT/r2i_tests.log 2>&1; tail -3 $OUT/r2i_tests.log
timeout 300 python tools/gemm_timeline4.py > $OUT/r2i_conv_split_timeline.log 2>&1; cat $OUT/r2i_conv_split_timeline.log
