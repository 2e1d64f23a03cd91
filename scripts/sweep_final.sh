#!/bin/bash
out=gpurun_out/sweep_final.txt
: > $out
cat > /tmp/time_decode.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
from pose_unsupervised_b200.core.inference import decode_heatmaps
hw = int(sys.argv[1]); N = 16384
g = torch.Generator(device='cuda').manual_seed(0)
hm = torch.rand((N, 17, hw, hw), generator=g, device='cuda')
c = torch.rand((N, 2), device='cuda', dtype=torch.float64) * 200 + 400
s = (torch.rand((N, 1), device='cuda', dtype=torch.float64) * 1.5 + 1.5).repeat(1, 2)
for _ in range(5): decode_heatmaps(hm, c, s, post_process=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(50): decode_heatmaps(hm, c, s, post_process=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
print('decode-only hw=%d %.4f ms %.0f GB/s' % (hw, ms, hm.numel() * 4 / ms / 1e6))
PY
for cfg in "2 2 2 2 2" "0 2 2 2 2" "2 2 3 2 2" "2 1 2 2 2" "2 4 2 2 2" "2 2 2 2 2"; do
  set -- $cfg
  export PB200_LIB=/tmp/libposeb200_fin_$1_$2_$3_$4_$5.so
  export PB200_NVCC_EXTRA="-DPB_PIPE_EPILOGUE=$1 -DPB_CLAIM_BATCH=$2 -DPB_STAGES=$3 -DPB_FUSED_MIN_BLOCKS=$4 -DPB_DECODE_MIN_BLOCKS=$5"
  python -m pose_unsupervised_b200.build --force > /dev/null 2>&1 || { echo "cfg $cfg BUILD FAILED" >> $out; continue; }
  for hw in 64 96; do
    r=$(timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-graph --hw $hw 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused kernel_ms %.4f %.0f GB/s' % (d['roofline']['kernel_ms'], d['roofline']['achieved']))" 2>&1)
    d=$(timeout 120 python /tmp/time_decode.py $hw 2>&1 | tail -1)
    echo "pipe=$1 batch=$2 stages=$3 fused_minblk=$4 decode_minblk=$5 hw=$hw : $r | $d" >> $out
  done
done
cat $out
