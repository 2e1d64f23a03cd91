#!/bin/bash
# RPSM level-0 strategy sweep (run on a GPU box): neighbourhood enumeration up to +-reach bins, sorted walk beyond.
out=gpurun_out/sweep_rpsm.txt
: > $out
for r in 2 3 4 5 16; do
  export PB200_LIB=/tmp/libposeb200_rpsm_$r.so
  export PB200_NVCC_EXTRA="-DPB_RPSM_ENUM_REACH=$r"
  python -m pose_unsupervised_b200.build --force > /dev/null 2>&1 || { echo "reach $r BUILD FAILED" >> $out; continue; }
  echo "reach=$r $(timeout 200 python bench.py --workload rpsm --steps 3 --frames 592 2>&1 | tail -1)" >> $out
done
cat $out
