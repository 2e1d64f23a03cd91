"""A stand-in for dataset.multiview_h36m (lib/dataset/multiview_h36m_compatible.py) with the record
layout the reference's evaluation scripts read: ``db[i]['camera' | 'joints_2d' | 'joints_3d' |
'center' | 'scale' | 'subject']`` and ``grouping`` = list of 4 db indices per frame.  No H36M data
ships with the reference, so frames are synthetic (seeded)."""
import numpy as np

from pose_unsupervised_b200.utils import synth


class SyntheticMultiViewH36M(object):
    def __init__(self, nframes=256, nviews=4, njoints=17, seed=0, noise_px=0.0):
        rng = np.random.default_rng(seed)
        self.rigs = synth.camera_table(7, nviews, seed=seed)
        self.poses = synth.random_poses(nframes, seed=seed + 1, njoints=njoints)
        self.subject = rng.integers(0, 7, nframes)
        self.db, self.grouping = [], []
        for f in range(nframes):
            items = []
            for v in range(nviews):
                cam = self.rigs[self.subject[f]][v]
                xy = synth.project_plumb_bob_numpy(self.poses[f], cam) + rng.normal(0, noise_px, (njoints, 2))
                lo, hi = xy.min(0), xy.max(0)
                s = float(max(hi - lo)) * 1.25 / 200.0
                cam_xyz = (cam['R'] @ (self.poses[f].T - cam['T'])).T
                self.db.append({'camera': cam, 'joints_2d': xy, 'joints_3d': cam_xyz,
                                'center': 0.5 * (lo + hi), 'scale': np.array([s, s]),
                                'subject': int(self.subject[f])})
                items.append(len(self.db) - 1)
            self.grouping.append(items)

    def heatmaps(self, hm_size=64, sigma=2.0, noise=0.02, seed=0):
        """[N, J, h, w] float32 network-like heatmaps rendered at the 2D joints of every row."""
        rng = np.random.default_rng(seed)
        n, j = len(self.db), self.db[0]['joints_2d'].shape[0]
        ys, xs = np.mgrid[0:hm_size, 0:hm_size].astype(np.float64)
        out = np.empty((n, j, hm_size, hm_size), dtype=np.float32)
        for i, rec in enumerate(self.db):
            t = synth.crop_affine_numpy(rec['center'], rec['scale'][0], hm_size, hm_size)
            p = rec['joints_2d'] @ t[:, :2].T + t[:, 2]
            for k in range(j):
                g = np.exp(-((xs - p[k, 0]) ** 2 + (ys - p[k, 1]) ** 2) / (2 * sigma ** 2))
                out[i, k] = (g + noise * rng.random((hm_size, hm_size))).astype(np.float32)
        return out
