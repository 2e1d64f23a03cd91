#!/usr/bin/env python
"""Counterpart of the reference's run/test/test_triangulate.py (:48-102) on a synthetic dataset:
heatmaps -> decode -> triangulate_poses -> MPJPE mean / std / max, through the reference's own
import lines.  `--heatmap` mimics the h5 hand-off (decode from heatmaps); without it the ground
truth 2D joints are triangulated, like the reference's flag_test_gt branch.

    python run/test/test_triangulate.py [--frames 256] [--heatmap] [--no-distortion]
"""
import argparse
import types

import numpy as np

import _init_paths  # noqa: F401
from core.inference import get_final_preds
from multiviews.cameras import camera_to_world_frame
from multiviews.triangulate import triangulate_poses

from synthetic_dataset import SyntheticMultiViewH36M


def main():
    ap = argparse.ArgumentParser(description='Triangulate multi-view 2d poses (B200 path, synthetic data)')
    ap.add_argument('--frames', type=int, default=256)
    ap.add_argument('--heatmap', action='store_true', help='decode 2D from synthetic heatmaps first')
    ap.add_argument('--no-distortion', action='store_true')
    args = ap.parse_args()
    test_dataset = SyntheticMultiViewH36M(args.frames)
    pred2d, cameras, gt3d = [], [], []
    for items in test_dataset.grouping:
        for item in items:
            cameras.append(test_dataset.db[item]['camera'])
            pred2d.append(test_dataset.db[item]['joints_2d'])
        gt = test_dataset.db[items[-1]]['joints_3d']
        gt3d.append(camera_to_world_frame(gt, cameras[-1]['R'], cameras[-1]['T']))
    pred2d, gt3d = np.array(pred2d), np.array(gt3d)
    if args.heatmap:
        config = types.SimpleNamespace(TEST=types.SimpleNamespace(POST_PROCESS=True))
        center = np.array([r['center'] for r in test_dataset.db])
        scale = np.array([r['scale'] for r in test_dataset.db])
        preds, maxvals = get_final_preds(config, test_dataset.heatmaps(), center, scale)
        pred2d = preds[:, :, :2]
    pred3d = triangulate_poses(cameras, pred2d, joints_vis=None, no_distortion=args.no_distortion)
    assert len(gt3d) == len(pred3d)
    norm = np.linalg.norm(pred3d - gt3d, axis=2)
    print('Mean Error:', np.mean(norm))
    print('Std Error:', np.std(norm))
    print('Max Error:', np.amax(norm))
    print('Larger than Mean+Std Error: {:.1%}'.format(np.sum(norm > np.mean(norm) + np.std(norm)) / norm.size))


if __name__ == '__main__':
    main()
