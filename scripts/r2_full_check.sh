#!/bin/bash
# round-2 single-GPU check: all GPU tests, smoke, the default bench line, launch list, ncu captures.
set -u
OUT=gpurun_out/r2d; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.log 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/bench.err
timeout 300 python bench.py --workload pseudo --steps 5 > $OUT/pseudo.log 2>&1; echo "pseudo rc=$?"
# launch list of the same default command (kernel shares of the step), then full captures of the top kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-pageable > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_tma -c 1 -o $OUT/prof_decode_tma \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-pageable --no-graph > $OUT/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rpsm_onchip -c 1 -o $OUT/prof_rpsm_onchip \
    python bench.py --workload rpsm --steps 1 --frames 592 --no-cpu-baseline > $OUT/ncu_rpsm.log 2>&1; echo "ncu rpsm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ransac_compact -c 2 -o $OUT/prof_ransac \
    python bench.py --workload pseudo --steps 1 > $OUT/ncu_ransac.log 2>&1; echo "ncu ransac rc=$?"
