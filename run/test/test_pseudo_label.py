#!/usr/bin/env python
"""Counterpart of the reference's run/test/test_pseudo_label.py (:142-258) on a synthetic dataset:
confidence threshold -> (RANSAC) -> reproject_poses -> pseudo labels + visibility, PCKh-style score
against the ground truth, written as .npz with the reference's dataset names (h5py is not available).

    python run/test/test_pseudo_label.py [--frames 512] [--no-ransac] [--out pseudo_label.npz]
"""
import argparse
import types

import numpy as np

import _init_paths  # noqa: F401
from multiviews.triangulate import ransac, reproject_poses

from synthetic_dataset import SyntheticMultiViewH36M


def pckh(pred, gt, vis, head_size, thr=0.5):
    d = np.linalg.norm(pred - gt, axis=2) / head_size[:, None]
    ok = (d <= thr) & (vis > 0)
    return ok.sum() / max(1, (vis > 0).sum())


def main():
    ap = argparse.ArgumentParser(description='Test pseudo labels (B200 path, synthetic data)')
    ap.add_argument('--frames', type=int, default=512)
    ap.add_argument('--no-ransac', action='store_true')
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    config = types.SimpleNamespace(
        DATASET=types.SimpleNamespace(NO_DISTORTION=False),
        PSEUDO_LABEL=types.SimpleNamespace(REPROJ_THRE=10, NUM_INLIERS=3, USE_REPROJ=True))
    ds = SyntheticMultiViewH36M(args.frames, noise_px=2.0, seed=3)
    rng = np.random.default_rng(4)
    gt2d = np.array([r['joints_2d'] for r in ds.db])
    pred2d = gt2d.copy()
    bad = rng.random(pred2d.shape[:2]) < 0.1
    pred2d[bad] += rng.normal(0, 50, (int(bad.sum()), 2))
    confidence = rng.uniform(0.04, 1.12, pred2d.shape[:2])
    cameras = [r['camera'] for r in ds.db]
    head = np.array([np.linalg.norm(r['joints_2d'][9] - r['joints_2d'][10]) for r in ds.db]) + 1e-6
    for conf_thre in [0.6, 0.7, 0.8, 0.9]:
        joints_vis = (confidence > conf_thre).astype(np.float64)
        before = pckh(pred2d, gt2d, joints_vis, head)
        if not args.no_ransac:
            joints_vis = ransac(pred2d, cameras, joints_vis, config)
        proj2d, joints_vis = reproject_poses(pred2d, cameras, joints_vis, config.DATASET.NO_DISTORTION)
        after = pckh(proj2d, gt2d, joints_vis, head)
        print('conf>%.1f  PCKh@0.5 before %.3f  after %.3f  visible ratio %.3f'
              % (conf_thre, before, after, joints_vis.mean()))
    if args.out:
        np.savez_compressed(args.out, pseudo_2d=proj2d, joints_vis=joints_vis)
        print('wrote', args.out)


if __name__ == '__main__':
    main()
