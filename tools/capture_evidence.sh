#!/bin/bash
# Round evidence: bench line, ncu launch list of the bench command, per-launch DRAM traffic of one step (C2 and C4), and
# full captures of the dominant kernels. Run on the GPU box: bash tools/capture_evidence.sh <tag>
# Numbers printed under ncu are never bench values.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 400 $OUT/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_step_traffic.csv \
  python tools/profile_step.py 3 C2 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_c4_step_traffic.csv \
  python tools/profile_step.py 3 C4 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_orient_step_traffic.csv \
  -k regex:"render_" python tools/profile_step.py 3 C2 orient > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"render_field_bwd|render_nerf_fwd2|render_composite_bwd|flash_attn" -c 8 \
  -o $OUT/${TAG}_full_misc -f python tools/profile_step.py 3 C2 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"render_orient" -c 4 -o $OUT/${TAG}_full_orient -f python tools/profile_step.py 3 C2 orient > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"hyper_field|volsdf|hypernet" -c 12 -o $OUT/${TAG}_full_c4 -f python tools/profile_step.py 3 C4 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:gemm_f16_kernel -s 70 -c 20 -o $OUT/${TAG}_full_gemm -f python tools/profile_step.py 3 C2 > /dev/null 2>&1
ls -la $OUT | tail -12
