#!/bin/bash
# Round evidence: bench line, ncu launch list of the bench command, per-launch DRAM traffic of one step (C2, C4, orientation),
# and full captures of the dominant kernels. Run on the GPU box: bash tools/capture_evidence.sh <tag>
# The .ncu-rep files stay in /tmp on the box (gpurun_out is limited to 64 MiB); their raw / source pages come back as CSV.
# Numbers printed under ncu are never bench values.
TAG=${1:-r2}
OUT=gpurun_out
REP=/tmp/ncu_reps
mkdir -p $OUT $REP
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 300 $OUT/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $OUT/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_step_traffic.csv \
  python tools/profile_step.py 3 C2 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_c4_step_traffic.csv \
  python tools/profile_step.py 3 C4 > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_orient_step_traffic.csv \
  -k regex:"render_" python tools/profile_step.py 3 C2 orient > /dev/null 2>&1
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $OUT/${TAG}_generator_step_traffic.csv \
  python tools/profile_generator.py > /dev/null 2>&1
full() {  # name, kernel regex, launches to capture, extra ncu args, workload args...
  local name=$1 rx=$2 cnt=$3 skip=$4; shift 4
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt \
    -o $REP/${TAG}_full_$name -f python tools/profile_step.py "$@" > /dev/null 2>&1
  ncu -i $REP/${TAG}_full_$name.ncu-rep --page raw --csv > $OUT/${TAG}_full_${name}_raw.csv 2>/dev/null
}
full misc "render_field_bwd|render_nerf_fwd2|render_composite_bwd|flash_attn" 8 0 3 C2
full orient "render_orient" 3 0 3 C2 orient
full c4 "hyper_field|volsdf_composite|hypernet" 7 0 3 C4
full gemm gemm_f16_kernel 10 70 3 C2
# the tf32 GEMM of the generator: the first launches of the forward (projections, a score product, P V)
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 4 -c 8 \
  -o $REP/${TAG}_full_gemm_tf32 -f python tools/profile_generator.py > /dev/null 2>&1
ncu -i $REP/${TAG}_full_gemm_tf32.ncu-rep --page raw --csv > $OUT/${TAG}_full_gemm_tf32_raw.csv 2>/dev/null
# source pages (per-instruction stall samples) of the two kernels DESIGN.md argues about
ncu -i $REP/${TAG}_full_gemm.ncu-rep --page source --csv 2>/dev/null | head -20000 > $OUT/${TAG}_full_gemm_source.csv
ncu -i $REP/${TAG}_full_misc.ncu-rep --page source --csv -k regex:render_field_bwd 2>/dev/null | head -12000 > $OUT/${TAG}_full_field_bwd_source.csv
du -sh $OUT; ls -la $OUT | tail -16
