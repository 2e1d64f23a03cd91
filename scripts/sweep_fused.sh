#!/bin/bash
# Where does the fused kernel lose time against the decode-only kernel?  (run on a GPU box)
out=gpurun_out/sweep_fused.txt
: > $out
for cfg in "0 1 0 3" "0 2 0 3" "0 4 0 3" "0 1 1 3" "0 4 1 3" "0 2 0 2" "0 4 0 2" "1 4 0 2" "1 4 0 3"; do
  set -- $cfg
  export PB200_LIB=/tmp/libposeb200_fu_$1_$2_$3_$4.so
  export PB200_NVCC_EXTRA="-DPB_PIPE_EPILOGUE=$1 -DPB_CLAIM_BATCH=$2 -DPB_SKIP_LIFT=$3 -DPB_STAGES=$4"
  python -m pose_unsupervised_b200.build --force > /dev/null 2>&1 || { echo "cfg $cfg BUILD FAILED" >> $out; continue; }
  for hw in 64 96; do
    r=$(timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-graph --hw $hw 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused kernel_ms %.4f %.0f GB/s' % (d['roofline']['kernel_ms'], d['roofline']['achieved']))" 2>&1)
    echo "pipe=$1 claim_batch=$2 skip_lift=$3 stages=$4 hw=$hw : $r" >> $out
  done
done
cat $out
