#!/usr/bin/env python
"""Counterpart of the reference's run/test/test_fund_mtx.py (:56-71) on a synthetic dataset: mean and
max algebraic epipolar residual |x_b^T F x_a| of 2D predictions over the 12 ordered view pairs, with F
taken exactly from the cameras (the reference fits F to data offline).

    python run/test/test_fund_mtx.py [--frames 256] [--noise 2.0]
"""
import argparse

import numpy as np

import _init_paths  # noqa: F401
from pose_unsupervised_b200.core.loss import FundamentalTable, epipolar_residuals

from synthetic_dataset import SyntheticMultiViewH36M


def main():
    ap = argparse.ArgumentParser(description='Epipolar residual of 2D predictions (B200 path, synthetic data)')
    ap.add_argument('--frames', type=int, default=256)
    ap.add_argument('--noise', type=float, default=2.0)
    args = ap.parse_args()
    ds = SyntheticMultiViewH36M(args.frames, noise_px=args.noise, seed=5)
    pred2d = np.array([r['joints_2d'] for r in ds.db])
    subjects = np.array([ds.db[items[0]]['subject'] for items in ds.grouping])
    table = FundamentalTable.from_cameras({s: ds.rigs[s] for s in range(len(ds.rigs))})
    res = epipolar_residuals(pred2d, subjects, table)
    print('mean: {}'.format(np.mean(res)))
    print('max: {}'.format(np.amax(res)))


if __name__ == '__main__':
    main()
