#!/bin/bash
# Round evidence: bench line, ncu launch list of the bench command, per-launch DRAM traffic of one step, and full
# captures of the dominant kernels. Run on the GPU box: bash tools/capture_evidence.sh <tag>
# Numbers printed under ncu are never bench values.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $OUT/${TAG}_step_traffic.csv python tools/profile_step.py 3 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"render_field_bwd|render_nerf_fwd2|render_composite_bwd|flash_attn" -c 8 \
  -o $OUT/${TAG}_full_misc -f python tools/profile_step.py 3 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:gemm_f16_kernel -s 70 -c 8 -o $OUT/${TAG}_full_gemm -f python tools/profile_step.py 3 > /dev/null 2>&1
ls -la $OUT | tail -8
