#!/usr/bin/env python
"""Counterpart of the reference's run/test/test_rpsm.py (:129-151) on synthetic frames: per frame
rpsm(cameras, hms, boxes, grid_center, limb_length, pairwise, config) -> MPJPE, plus the batched call.

    python run/test/test_rpsm.py [--frames 8]
"""
import argparse
import types

import numpy as np

import _init_paths  # noqa: F401
from multiviews.body import HumanBody
from multiviews.pictorial import PairwiseTable, rpsm, rpsm_batch

from pose_unsupervised_b200.utils import synth


def compute_limb_length(body, pose):
    return {(n['idx'], c): float(np.linalg.norm(pose[n['idx']] - pose[c]))
            for n in body.skeleton for c in n['children']}


def main():
    ap = argparse.ArgumentParser(description='Test Recursive Pictorial Structure Model (B200 path, synthetic data)')
    ap.add_argument('--frames', type=int, default=8)
    args = ap.parse_args()
    config = types.SimpleNamespace(
        NETWORK=types.SimpleNamespace(IMAGE_SIZE=np.array([256, 256]), HEATMAP_SIZE=np.array([64, 64])),
        PICT_STRUCT=types.SimpleNamespace(FIRST_NBINS=16, RECUR_NBINS=2, RECUR_DEPTH=10, GRID_SIZE=2000,
                                          LIMB_LENGTH_TOLERANCE=150))
    body = HumanBody()
    poses = synth.random_poses(args.frames, seed=1, njoints=16)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99, njoints=16)]))
           for e in body.edges()}
    pairwise = PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    res, frames = [], []
    for f in range(args.frames):
        cameras = synth.camera_ring(4, seed=10 + f)
        boxes = synth.crop_box(cameras, poses[f])
        hms = synth.gaussian_heatmaps(cameras, boxes, poses[f], 64, 256, 2.0, 0.02, seed=f)
        limb_length = compute_limb_length(body, poses[f])
        grid_center = poses[f][body.root_idx]
        pose = rpsm(cameras, hms, boxes, grid_center, limb_length, pairwise, config)
        res.append(np.mean(np.sqrt(np.sum((pose - poses[f]) ** 2, axis=1))))
        frames.append((cameras, hms, boxes, grid_center, limb_length))
        print('%d:%.2f' % (f, res[-1]))
    print('MPJPE: ', np.mean(res))
    cams = [c for fr in frames for c in fr[0]]
    batch = rpsm_batch(cams, np.array([fr[1] for fr in frames]),
                       np.array([b['center'] for fr in frames for b in fr[2]]),
                       np.array([b['scale'] for fr in frames for b in fr[2]]),
                       np.array([fr[3] for fr in frames]),
                       np.array([[fr[4][e] for e in body.edges()] for fr in frames]), pairwise, config, body)
    print('MPJPE (batched call): ', np.mean(np.linalg.norm(batch - poses, axis=2)))


if __name__ == '__main__':
    main()
