#!/bin/bash
# ncu capture of the on-chip RPSM kernel (one launch, 592 frames) -> gpurun_out/r2n/
set -u
OUT=gpurun_out/r2n; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rpsm_onchip -c 1 -f -o $OUT/prof_rpsm_onchip python bench.py --workload rpsm --steps 1 --frames 592 --no-cpu-baseline > $OUT/ncu_rpsm.log 2>&1; echo "ncu rc=$?"
