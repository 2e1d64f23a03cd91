#!/bin/bash
# static vs dynamic decode schedule under the overlapped exchange, N GPUs
set -u
N=${1:-2}; OUT=gpurun_out/r2g; mkdir -p $OUT
for mode in ${MODES:-static dynamic}; do
  T0=$SECONDS
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N --steps 200 --warmup 5 --no-e2e --decode-schedule $mode >> $OUT/${mode}_${N}gpu.log 2>> $OUT/${mode}_${N}gpu.err
  echo "$mode N=$N rc=$? wall=$((SECONDS - T0))s"
done
