"""CUDA-event timings of the secondary kernels (run on a GPU box); one JSON line per kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pose_unsupervised_b200.core.inference import decode_heatmaps, decode_heatmaps_flip   # noqa: E402
from pose_unsupervised_b200.utils.transforms import generate_integral_preds_2d_th          # noqa: E402


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


N, J, HW = 8192, 17, 64
g = torch.Generator(device='cuda').manual_seed(0)
hm = torch.rand((N, J, HW, HW), generator=g, device='cuda')
hf = torch.rand((N, J, HW, HW), generator=g, device='cuda')
c = torch.rand((N, 2), device='cuda', dtype=torch.float64) * 200 + 400
s = (torch.rand((N, 1), device='cuda', dtype=torch.float64) * 1.5 + 1.5).repeat(1, 2)
nbytes = hm.numel() * 4
pairs = [[1, 4], [2, 5], [3, 6], [11, 14], [12, 15], [13, 16]]

ms = timeit(lambda: decode_heatmaps(hm, c, s, post_process=True))
print(json.dumps({'kernel': 'decode_tma_kernel (get_final_preds)', 'rows': N, 'ms': ms, 'GBps': nbytes / ms / 1e6}))
ms = timeit(lambda: decode_heatmaps(hm))
print(json.dumps({'kernel': 'decode_tma_kernel (get_max_preds)', 'rows': N, 'ms': ms, 'GBps': nbytes / ms / 1e6}))
ms = timeit(lambda: decode_heatmaps_flip(hm, hf, pairs, True, c, s, True))
print(json.dumps({'kernel': 'decode_flip_kernel (2 reads + 1 write per element)', 'rows': N, 'ms': ms,
                  'GBps': 3 * nbytes / ms / 1e6}))


def torch_flip_recipe():
    order = list(range(J))
    for a, b in pairs:
        order[a], order[b] = b, a
    fb = torch.index_select(torch.flip(hf, dims=[3]), 1, torch.tensor(order, device='cuda'))
    fb[:, :, :, 1:] = fb.clone()[:, :, :, 0:-1]
    return decode_heatmaps((hm + fb) * 0.5, c, s, post_process=True)


ms = timeit(torch_flip_recipe, n=10)
print(json.dumps({'kernel': 'torch flip/index_select/clone/add/mul + decode (validate() recipe)', 'rows': N, 'ms': ms}))
h2 = hm[:4096].clone().requires_grad_(True)
ms = timeit(lambda: generate_integral_preds_2d_th(h2))
print(json.dumps({'kernel': 'softargmax_fwd_kernel', 'rows': 4096, 'ms': ms, 'GBps': h2.numel() * 4 / ms / 1e6}))
xy = generate_integral_preds_2d_th(h2)
gxy = torch.ones_like(xy)
ms = timeit(lambda: torch.autograd.grad(xy, h2, gxy, retain_graph=True))
print(json.dumps({'kernel': 'softargmax_bwd_kernel (1 read + 1 write)', 'rows': 4096, 'ms': ms,
                  'GBps': 2 * h2.numel() * 4 / ms / 1e6}))


def torch_softargmax(h):
    n, j, hh, ww = h.shape
    p = torch.nn.functional.softmax((h * 100).view(n, j, -1), dim=-1).view(n, j, hh, ww)
    xs = torch.arange(ww, dtype=torch.float32, device=h.device)
    ys = torch.arange(hh, dtype=torch.float32, device=h.device)
    return torch.stack([(p.sum(dim=2) * xs.view(1, 1, -1)).sum(dim=2), (p.sum(dim=3) * ys.view(1, 1, -1)).sum(dim=2)], dim=2)


ms = timeit(lambda: torch_softargmax(h2), n=10)
print(json.dumps({'kernel': 'torch soft-argmax recipe (lib/utils/transforms.py:149-171), forward', 'rows': 4096, 'ms': ms}))
