#!/bin/bash
# quick GPU check of an RPSM change: parity tests, then throughput at the two batch sizes quoted in DESIGN.md
set -u
OUT=gpurun_out/r2q; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_rpsm.py -x -q > $OUT/pytest_rpsm.log 2>&1; echo "pytest_rpsm rc=$?" | tee -a $OUT/pytest_rpsm.log
tail -3 $OUT/pytest_rpsm.log
timeout 300 python bench.py --workload rpsm --steps 5 --no-cpu-baseline > $OUT/rpsm.log 2>&1; echo "rpsm rc=$?"
timeout 300 python bench.py --workload rpsm --steps 5 --frames 2368 --no-cpu-baseline > $OUT/rpsm_2368.log 2>&1; echo "rpsm2368 rc=$?"
tail -q -n 1 $OUT/rpsm.log $OUT/rpsm_2368.log
