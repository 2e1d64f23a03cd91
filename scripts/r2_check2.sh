#!/bin/bash
set -u
OUT=gpurun_out/r2i; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python bench.py --workload rpsm --steps 5 --frames 2368 --no-cpu-baseline > $OUT/rpsm_2368.log 2>&1; echo "rpsm2368 rc=$?"
timeout 600 python bench.py > $OUT/bench.log 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rpsm_onchip -c 1 -o $OUT/prof_rpsm_onchip python bench.py --workload rpsm --steps 1 --frames 592 --no-cpu-baseline > $OUT/ncu_rpsm.log 2>&1; echo "ncu rc=$?"
