#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_render_gpu.py tests/test_guidance_gpu.py -x -q -s -k "orient or z_variance or survives" > $OUT/r2b_orient.log 2>&1
tail -30 $OUT/r2b_orient.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_render_gpu.py -x -q -k "orientation" > $OUT/r2b_sanitizer.log 2>&1
tail -8 $OUT/r2b_sanitizer.log
timeout 300 python tools/gemm_shapes.py default > $OUT/r2b_gemm_default.log 2>&1; tail -30 $OUT/r2b_gemm_default.log
SDB_GEMM_BN=128 timeout 300 python tools/gemm_shapes.py bn128 > $OUT/r2b_gemm_bn128.log 2>&1; tail -30 $OUT/r2b_gemm_bn128.log
SDB_GEMM_BN=64 timeout 300 python tools/gemm_shapes.py bn64 > $OUT/r2b_gemm_bn64.log 2>&1
SDB_GEMM_DEEP=all timeout 300 python tools/gemm_shapes.py deepall > $OUT/r2b_gemm_deepall.log 2>&1
SDB_GEMM_DEEP=none timeout 300 python tools/gemm_shapes.py deepnone > $OUT/r2b_gemm_deepnone.log 2>&1
