#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_kernel -s 2 -c 1 -o $OUT/r2q_gemm_20480x2560x320 -f python tools/ncu_gemm_one.py 20480 2560 320 > $OUT/r2q_ncu.log 2>&1
tail -3 $OUT/r2q_ncu.log; ls -la $OUT/*.ncu-rep
