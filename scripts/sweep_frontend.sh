#!/bin/bash
# Front-end code-shape sweep (run on a GPU box): epilogue pipelining x LDS hoisting x stages,
# fused kernel and decode-only kernel, 64x64 and 96x96.
out=gpurun_out/sweep_frontend.txt
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
for cfg in "0 0 3" "1 0 3" "0 1 3" "1 1 3" "1 0 2" "1 1 2" "0 0 4"; do
  set -- $cfg
  w=8; if [ "$3" = "4" ]; then w=6; fi
  export PB200_LIB=/tmp/libposeb200_fe_$1_$2_$3.so
  export PB200_NVCC_EXTRA="-DPB_PIPE_EPILOGUE=$1 -DPB_SCAN_HOIST=$2 -DPB_STAGES=$3 -DPB_FUSED_WARPS=$w"
  python -m pose_unsupervised_b200.build --force > /dev/null 2>&1 || { echo "cfg $cfg BUILD FAILED" >> $out; continue; }
  for hw in 64 96; do
    r=$(timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-graph --hw $hw 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused kernel_ms %.4f %.0f GB/s' % (d['roofline']['kernel_ms'], d['roofline']['achieved']))" 2>&1)
    d=$(timeout 120 python /tmp/time_decode.py $hw 2>&1 | tail -1)
    echo "pipe=$1 hoist=$2 stages=$3 warps=$w hw=$hw : $r | $d" >> $out
  done
done
cat $out
