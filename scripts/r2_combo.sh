#!/bin/bash
# decode schedule A/B (1 GPU) + RPSM parity and speed
set -u
bash scripts/r2_decode_ab.sh
OUT=gpurun_out/r2c; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_rpsm.py -x -q > $OUT/pytest_rpsm.log 2>&1; echo "pytest_rpsm rc=$?" | tee -a $OUT/pytest_rpsm.log
timeout 300 python bench.py --workload rpsm --steps 5 --frames 2368 --no-cpu-baseline > $OUT/rpsm_2368.log 2>&1; echo "rpsm2368 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rpsm_onchip -c 1 -o $OUT/prof_rpsm_onchip python bench.py --workload rpsm --steps 1 --frames 592 --no-cpu-baseline > $OUT/ncu_rpsm.log 2>&1; echo "ncu rc=$?"
