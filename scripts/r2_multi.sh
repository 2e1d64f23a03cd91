#!/bin/bash
# round-2 multi-GPU check on N GPUs of one box: default bench (exit code + wall time), pseudo-label passes.
set -u
N=${1:-2}; OUT=gpurun_out/r2e; mkdir -p $OUT
run() {  # name, args...
  local name=$1; shift
  local T0=$SECONDS
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N "$@" > $OUT/${name}_${N}gpu.log 2> $OUT/${name}_${N}gpu.err
  echo "$name N=$N rc=$? wall=$((SECONDS - T0))s" | tee -a $OUT/summary_${N}gpu.txt
}
run bench --steps 200 --warmup 5 --no-pageable
run bench20 --steps 20 --warmup 5 --no-pageable
if [ "${2:-}" = "all" ]; then
  run pseudo --workload pseudo --steps 5
  run pseudohm --workload pseudo-hm --steps 3
  run reference --impl reference --steps 2 --warmup 1
fi
if [ "${2:-}" = "all" ] || [ "${2:-}" = "sweep" ]; then
  run sweep_v8_hw64 --views 8 --hw 64 --steps 48 --warmup 3 --no-e2e --no-cpu-baseline
  run sweep_v4_hw96 --views 4 --hw 96 --steps 48 --warmup 3 --no-e2e --no-cpu-baseline
  run sweep_v2_hw64 --views 2 --hw 64 --steps 48 --warmup 3 --no-e2e --no-cpu-baseline
fi
grep -c "death signal" $OUT/*_${N}gpu.err | tee -a $OUT/summary_${N}gpu.txt
