#!/bin/bash
# Tuning sweep of the fused lift kernel's TMA ring (run on a GPU box):
#   chunk floats x stages x warps per block x min blocks per SM
# Each configuration is built into its own .so (PB200_LIB) and timed with bench.py.
set -u
out=gpurun_out/sweep_lift.txt
: > $out
python - <<'PY' >> $out 2>&1
import torch, time
x = torch.rand((4096*4, 17, 64, 64), device='cuda')
for name, fn in [('torch.sum', lambda: x.sum()), ('torch.amax(dim=(2,3))', lambda: x.amax(dim=(2, 3))),
                 ('torch.argmax(flat)', lambda: x.view(4096*4, 17, -1).argmax(dim=2))]:
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print('calib %-24s %.4f ms  %.0f GB/s' % (name, ms, x.numel() * 4 / ms / 1e6))
PY
for cfg in "1024 3 8 2" "1024 2 12 2" "1024 2 8 3" "1024 4 6 2" "2048 2 6 2" "512 4 12 2" "512 3 16 2" "1024 3 16 1" "1024 2 10 2"; do
  set -- $cfg
  export PB200_LIB=/tmp/libposeb200_$1_$2_$3_$4.so
  export PB200_NVCC_EXTRA="-DPB_CHUNK_FLOATS=$1 -DPB_STAGES=$2 -DPB_FUSED_WARPS=$3 -DPB_FUSED_MIN_BLOCKS=$4"
  python -m pose_unsupervised_b200.build --force > /dev/null 2>&1 || { echo "cfg $cfg BUILD FAILED" >> $out; continue; }
  res=$(timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kernel_ms %.4f frac %.3f value %.3e' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['value']))" 2>&1)
  echo "cfg chunk=$1 stages=$2 warps=$3 minblk=$4 : $res" >> $out
done
cat $out
