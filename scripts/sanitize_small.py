"""Smallest invocation of every kernel, for compute-sanitizer (memcheck / racecheck) on a GPU box:

    python scripts/sanitize_small.py && compute-sanitizer --tool memcheck python scripts/sanitize_small.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pose_unsupervised_b200 import _lib                                      # noqa: E402
from pose_unsupervised_b200.core.inference import decode_heatmaps            # noqa: E402
from pose_unsupervised_b200.core.loss import epipolar_residuals              # noqa: E402
from pose_unsupervised_b200.multiviews import cameras, pictorial, triangulate  # noqa: E402
from pose_unsupervised_b200.multiviews.body import HumanBody                 # noqa: E402
from pose_unsupervised_b200.utils import synth                               # noqa: E402
from tests.util import pseudo_config, rpsm_config                            # noqa: E402

rng = np.random.default_rng(0)
B, V, J = 6, 4, 17
rig = synth.camera_ring(V, seed=0)
cams = [rig[v] for _ in range(B) for v in range(V)]
center = rng.uniform(400, 600, (B * V, 2))
scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
for hw in (64, 80, 17):
    hm = rng.random((B * V, J, hw, hw), dtype=np.float32)
    hm[1, 2, 3, 4] = np.nan
    decode_heatmaps(hm)
    decode_heatmaps(hm, center, scale, post_process=True, return_idx=True)
    for schedule in (0, 1):
        _lib.call('pb200_set_tuning', _lib.TUNE_DECODE_SCHEDULE, schedule)
        triangulate.lift_heatmaps(hm, center, scale, cams, conf_thre=0.5, return_idx=True, return_proj=True)
poses = synth.random_poses(B, seed=1)
obs, cams2 = synth.multiview_observations(poses, [rig], [0] * B, noise_px=2.0, outlier_frac=0.1, seed=2)
vis = (rng.random(obs.shape[:2]) > 0.2).astype(np.float64)
triangulate.triangulate_poses(cams2, obs, vis)
triangulate.reproject_poses(obs, cams2, vis, return_points=True)
triangulate.ransac(obs, cams2, vis, pseudo_config())
st = triangulate.mpjpe_stats(poses, poses + 1.0)
cameras.project_pose(poses[0], rig[0])
cameras.camera_to_world_frame(poses[0], rig[0]['R'], rig[0]['T'])
F = {(0, a, b): np.eye(3) for a in range(4) for b in range(4) if a != b}
epipolar_residuals(obs, [0] * B, F, return_sum=True)
body = HumanBody.h36m17()
edges = body.edges()
cfg = rpsm_config(depth=2)
avg = {e: float(np.linalg.norm(synth.H36M17_REST[e[0]] - synth.H36M17_REST[e[1]])) for e in edges}
table = pictorial.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
assert table.offset_only
boxes = synth.crop_box(rig, poses[0])
hm = synth.gaussian_heatmaps(rig, boxes, poses[0], 64, 256, 2.0, 0.02, seed=0)
limb = synth.limb_lengths(poses[0], edges)
for use_lut, onchip in ((True, True), (True, False), (False, False)):
    pictorial.rpsm_batch(rig, hm[None], np.array([b['center'] for b in boxes]),
                         np.array([b['scale'] for b in boxes]), poses[0][:1], np.array([[limb[e] for e in edges]]),
                         table, cfg, body, return_trace=True, use_lut=use_lut, onchip=onchip)
torch.cuda.synchronize()
print('sanitize_small ok')
