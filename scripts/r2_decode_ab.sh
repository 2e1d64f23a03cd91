#!/bin/bash
# A/B of the decode map schedule on one GPU: dynamic claims (default build) vs static striding.
set -u
OUT=gpurun_out/r2f; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_lift.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"
for i in 1 2; do
timeout 200 python bench.py --steps 200 --warmup 10 --no-secondary --no-cpu-baseline --no-e2e --decode-schedule dynamic > $OUT/dyn_$i.log 2>&1; echo "dyn rc=$?"
timeout 200 python bench.py --steps 200 --warmup 10 --no-secondary --no-cpu-baseline --no-e2e --decode-schedule static > $OUT/static_$i.log 2>&1; echo "static rc=$?"
done
