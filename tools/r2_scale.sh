#!/bin/bash
# bash tools/r2_scale.sh N [steps]: the driver's launch line for N GPUs (C2 line + the measured c4 block)
N=${1:-2}; K=${2:-10}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps $K --warmup 3 > $OUT/r2_scale_n$N.json 2> $OUT/r2_scale_n$N.err
tail -c 3500 $OUT/r2_scale_n$N.json; tail -5 $OUT/r2_scale_n$N.err | cut -c1-300
