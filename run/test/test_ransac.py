#!/usr/bin/env python
"""Counterpart of the reference's run/test/test_ransac.py (:60-121) on a synthetic dataset: RANSAC view
selection over noisy 2D joints with outliers, triangulation with the inlier mask, MPJPE and the
error histogram the reference prints.

    python run/test/test_ransac.py [--frames 256] [--outliers 0.1]
"""
import argparse
import types

import numpy as np

import _init_paths  # noqa: F401
from multiviews.cameras import camera_to_world_frame
from multiviews.triangulate import ransac, triangulate_poses

from synthetic_dataset import SyntheticMultiViewH36M


def main():
    ap = argparse.ArgumentParser(description='Triangulate with RANSAC view selection (B200 path, synthetic data)')
    ap.add_argument('--frames', type=int, default=256)
    ap.add_argument('--outliers', type=float, default=0.1)
    ap.add_argument('--no-distortion', action='store_true')
    args = ap.parse_args()
    config = types.SimpleNamespace(
        DATASET=types.SimpleNamespace(NO_DISTORTION=args.no_distortion),
        PSEUDO_LABEL=types.SimpleNamespace(REPROJ_THRE=10, NUM_INLIERS=3))
    ds = SyntheticMultiViewH36M(args.frames, noise_px=1.5, seed=7)
    rng = np.random.default_rng(8)
    pred2d, cameras, gt3d = [], [], []
    for items in ds.grouping:
        for item in items:
            cameras.append(ds.db[item]['camera'])
            pred2d.append(ds.db[item]['joints_2d'])
        gt = ds.db[items[-1]]['joints_3d']
        gt3d.append(camera_to_world_frame(gt, cameras[-1]['R'], cameras[-1]['T']))
    pred2d, gt3d = np.array(pred2d), np.array(gt3d)
    bad = rng.random(pred2d.shape[:2]) < args.outliers
    pred2d[bad] += rng.normal(0, 60, (int(bad.sum()), 2))

    def report(tag, pred3d, mask=None):
        norm = np.linalg.norm(pred3d - gt3d, axis=2)
        if mask is not None:
            norm = norm[mask]
        print('-- %s --' % tag)
        print('Mean Error:', np.mean(norm))
        print('Std Error:', np.std(norm))
        print('Max Error:', np.amax(norm))
        thre_list = [10, 20, 30, 40, 50, 100, 200, 500, 1000]
        print('| ' + ' | '.join(str(t) for t in thre_list) + ' |')
        print(''.join('| {:.1%} '.format(np.sum(norm < t) / norm.size) for t in thre_list) + '|')

    report('plain triangulation of all views', triangulate_poses(cameras, pred2d, None, args.no_distortion))
    joints_vis = np.ones(pred2d.shape[:2])
    joints_vis = ransac(camera_params=cameras, poses2d=pred2d, joints_vis=joints_vis, config=config)
    pred3d = triangulate_poses(cameras, pred2d, joints_vis=joints_vis, no_distortion=args.no_distortion)
    kept = joints_vis.reshape(len(gt3d), 4, -1).sum(axis=1) >= 2
    report('after RANSAC (joints with >= 2 inlier views: %.1f%%)' % (100 * kept.mean()), pred3d, kept)


if __name__ == '__main__':
    main()
