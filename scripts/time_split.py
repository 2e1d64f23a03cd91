"""Fused lift kernel against the two-kernel path (decode_tma_kernel + geometry_kernel) on the headline workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_side_inputs
from pose_unsupervised_b200 import _lib, runtime as rt
from pose_unsupervised_b200.core.inference import decode_heatmaps
from pose_unsupervised_b200.multiviews.cameras import CameraTable
from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps, reproject_poses, triangulate_poses

B, V, J, HW = 4096, 4, 17, 64
rigs, subj, pack, index, center, scale = make_side_inputs(B, 0)
table = CameraTable.from_arrays(pack, index)
g = torch.Generator(device='cuda').manual_seed(1)
hm = torch.rand((B * V, J, HW, HW), generator=g, device='cuda')
dc, ds = rt.to_device(center), rt.to_device(scale)
vis = torch.ones((B * V, J), dtype=torch.uint8, device='cuda')


def timeit(fn, n=100):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def fused():
    _lib.call('pb200_set_tuning', 1, 1)          # one persistent kernel, TMA ring front end
    try:
        return lift_heatmaps(hm, dc, ds, table, return_proj=True)
    finally:
        _lib.call('pb200_set_tuning', 1, 2)      # default: two kernels


def split():
    xy, mv = decode_heatmaps(hm, dc, ds, post_process=True)
    return reproject_poses(xy, table, vis, return_points=True)


def split_graph():
    pass


xy, mv = decode_heatmaps(hm, dc, ds, post_process=True)
print('fused           %.4f ms' % timeit(fused))
print('split (eager)   %.4f ms' % timeit(split))
print('  decode only   %.4f ms' % timeit(lambda: decode_heatmaps(hm, dc, ds, post_process=True)))
print('  reproject only %.4f ms' % timeit(lambda: reproject_poses(xy, table, vis, return_points=True)))
print('  triangulate only %.4f ms' % timeit(lambda: triangulate_poses(table, xy)))
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    split()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(gr):
    out = split()
print('split (graph)   %.4f ms' % timeit(gr.replay))
f = fused()
assert torch.equal(f.poses3d, out[2]) and torch.equal(f.proj2d.float(), out[0])
print('identical results')
