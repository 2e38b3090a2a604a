#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_amortized_gpu.py tests/test_fullsize_parity_gpu.py -x -q -s -k "hypernetwork or fullsize or 256 or step" > $OUT/r2g_tests.log 2>&1
grep -E "C2 256|C4 256|kept samples|comp_rgb|opacity|depth|passed|failed|Error|error|assert" $OUT/r2g_tests.log | tail -30
