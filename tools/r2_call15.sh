#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_nn_kernels_gpu.py -x -q 2>&1 | tail -2
timeout 300 python tools/gemm_shapes.py acc2 > $OUT/r2l_gemm_acc2.log 2>&1; tail -1 $OUT/r2l_gemm_acc2.log
SDB_GEMM_BN=128 timeout 300 python tools/gemm_shapes.py acc2bn128 > $OUT/r2l_gemm_acc2_bn128.log 2>&1; cat $OUT/r2l_gemm_acc2_bn128.log
SDB_GEMM_BN=64 timeout 300 python tools/gemm_shapes.py acc2bn64 > $OUT/r2l_gemm_acc2_bn64.log 2>&1; tail -1 $OUT/r2l_gemm_acc2_bn64.log
