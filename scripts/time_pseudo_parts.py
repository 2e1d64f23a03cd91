"""Break-down of the pseudo-label pass (1 M frames) into its kernels, CUDA events (run on a GPU box)."""
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pose_unsupervised_b200 import runtime as rt                                              # noqa: E402
from pose_unsupervised_b200.core.loss import FundamentalTable, epipolar_residuals            # noqa: E402
from pose_unsupervised_b200.multiviews.cameras import CameraTable, pack_camera               # noqa: E402
from pose_unsupervised_b200.multiviews.triangulate import ransac, reproject_poses, triangulate_poses  # noqa: E402
from pose_unsupervised_b200.utils import synth                                               # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rng = np.random.default_rng(0)
rigs = synth.camera_table(7, 4, seed=0)
pack = np.array([pack_camera(c) for rig in rigs for c in rig])
subj = rng.integers(0, 7, B)
table = CameraTable.from_arrays(pack, (subj[:, None] * 4 + np.arange(4)[None]).reshape(-1))
ftab = FundamentalTable.from_cameras({s: rigs[s] for s in range(7)})
slots = ftab.slots(subj)
obs = torch.rand((B * 4, 17, 2), device='cuda') * 600 + 200
conf = torch.rand((B * 4, 17), device='cuda') * 1.08 + 0.04
cfg = types.SimpleNamespace(DATASET=types.SimpleNamespace(NO_DISTORTION=False),
                            PSEUDO_LABEL=types.SimpleNamespace(REPROJ_THRE=10.0, NUM_INLIERS=3))
# consistent observations for a realistic inlier structure
poses = torch.from_numpy(synth.random_poses(1024, seed=5)[rng.integers(0, 1024, min(B, 200000))])


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


vis = conf > 0.7
allvis = torch.ones_like(vis)
for name, fn in [('triangulate_poses (all visible)', lambda: triangulate_poses(table, obs)),
                 ('reproject_poses (conf > 0.7)', lambda: reproject_poses(obs, table, vis, False, return_points=True)),
                 ('reproject_poses (all visible)', lambda: reproject_poses(obs, table, allvis, False, return_points=True)),
                 ('ransac (conf > 0.7)', lambda: ransac(obs, table, vis, cfg)),
                 ('ransac (all visible: 6 pairs per joint)', lambda: ransac(obs, table, allvis, cfg)),
                 ('epipolar_residuals', lambda: epipolar_residuals(obs, slots, ftab))]:
    ms = timeit(fn)
    print(json.dumps({'kernel': name, 'frames': B, 'ms': ms, 'Mframes_per_s': B / ms / 1e3}))
