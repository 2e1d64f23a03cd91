"""Oracle: per-joint multi-view triangulation, RANSAC view selection, reprojection.

TEST INFRASTRUCTURE, see oracle/__init__.py.  Control flow follows
lib/multiviews/triangulate.py:17-213 (visibility rules, zero fill for <2
views, ``itertools.combinations`` pair order, inlier tie-breaking); the three
pymvg calls are the restatements in oracle/pymvg_restated.py (PARITY UNPINNED
at that boundary).  The reference hard-codes ``nviews = 4``
(triangulate.py:70,114,183); here it is a keyword so the 2/4/8-view sweep of
BASELINE.json can be checked.  Rows are view-minor: row = frame*nviews + view.
"""
import itertools

import numpy as np

from .cameras import camera_fields
from .pymvg_restated import RestatedCamera, RestatedMultiCameraSystem


def _cam_name(j):
    return 'camera_{}'.format(j)


def build_multi_camera_system(cameras, no_distortion=False):
    """lib/multiviews/triangulate.py:17-40.  ``cameras``: list of (name, dict)."""
    rig = []
    for name, cam in cameras:
        R, T, f, c, k, p = camera_fields(cam, avg_f=False)
        f = np.asarray(f, dtype=np.float64).reshape(-1)      # numpy-2 adapter for :29-30
        c = np.asarray(c, dtype=np.float64).reshape(-1)
        k = np.asarray(k, dtype=np.float64).reshape(-1)
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        K = np.array([[f[0], 0, c[0]], [0, f[1], c[1]], [0, 0, 1]], dtype=float)
        dist = np.array([k[0], k[1], p[0], p[1], k[2]])
        t = -np.matmul(R, np.asarray(T, dtype=np.float64).reshape(3, 1))
        M = K.dot(np.concatenate((R, t), axis=1))
        rig.append(RestatedCamera.load_camera_from_M(
            M, name=name, distortion_coefficients=None if no_distortion else dist))
    return RestatedMultiCameraSystem(rig)


def _frame_rig(camera_params, i, nviews, no_distortion):
    return build_multi_camera_system(
        [(_cam_name(j), camera_params[i * nviews + j]) for j in range(nviews)], no_distortion)


def _visible_obs(poses2d, joints_vis, i, k, nviews):
    return [(_cam_name(j), poses2d[i * nviews + j, k, :])
            for j in range(nviews) if joints_vis[i * nviews + j, k]]


def triangulate_poses(camera_params, poses2d, joints_vis=None, no_distortion=False, nviews=4):
    """lib/multiviews/triangulate.py:57-99 -> poses3d [N/nviews, J, 3] float64."""
    njoints = poses2d.shape[1]
    nframes = len(camera_params) // nviews
    if joints_vis is not None:
        assert np.all(joints_vis.shape == poses2d.shape[:2])
    else:
        joints_vis = np.ones((poses2d.shape[0], poses2d.shape[1]))
    out = np.zeros((nframes, njoints, 3))
    for i in range(nframes):
        rig = _frame_rig(camera_params, i, nviews, no_distortion)
        for k in range(njoints):
            obs = _visible_obs(poses2d, joints_vis, i, k, nviews)
            if len(obs) < 2:
                continue                              # joint stays at zeros (:95-96)
            out[i, k, :] = rig.find3d(obs)
    return out


def ransac(poses2d, camera_params, joints_vis, reproj_thre, num_inliers,
           no_distortion=False, nviews=4):
    """lib/multiviews/triangulate.py:102-166.

    ``reproj_thre`` / ``num_inliers`` / ``no_distortion`` are
    ``config.PSEUDO_LABEL.REPROJ_THRE`` / ``.NUM_INLIERS`` /
    ``config.DATASET.NO_DISTORTION``.
    """
    njoints = poses2d.shape[1]
    nframes = len(camera_params) // nviews
    res_vis = np.zeros_like(joints_vis)
    for i in range(nframes):
        rig = _frame_rig(camera_params, i, nviews, no_distortion)
        for k in range(njoints):
            obs = _visible_obs(poses2d, joints_vis, i, k, nviews)
            if len(obs) < 2:
                continue
            best_views, best_err = [], 10000
            for pair in itertools.combinations(obs, 2):
                X = rig.find3d(list(pair))
                views, err_sum = [], 0
                for j in range(nviews):
                    e = np.linalg.norm(rig.find2d(_cam_name(j), X) - poses2d[i * nviews + j, k, :])
                    if e < reproj_thre:
                        views.append(j)
                        err_sum += e
                if len(views) < num_inliers:
                    continue
                mean_err = err_sum / len(views)
                if len(views) > len(best_views) or \
                        (len(views) == len(best_views) and mean_err < best_err):
                    best_views, best_err = views, mean_err
            for j in best_views:
                res_vis[i * nviews + j, k] = 1
    return res_vis


def ransac_pairs(poses2d, camera_params, joints_vis, frame, joint, no_distortion=False, nviews=4):
    """Diagnostics for one (frame, joint) of ``ransac``: for every pair of visible views, in
    itertools.combinations order, (pair, reprojection error in each of the nviews views).  Tests use it
    to show that a selection that differs from the GPU's sits on a rounding-level near-tie."""
    rig = _frame_rig(camera_params, frame, nviews, no_distortion)
    obs = _visible_obs(poses2d, joints_vis, frame, joint, nviews)
    out = []
    for pair in itertools.combinations(obs, 2):
        X = rig.find3d(list(pair))
        errs = [float(np.linalg.norm(rig.find2d(_cam_name(j), X) - poses2d[frame * nviews + j, joint, :]))
                for j in range(nviews)]
        out.append(((pair[0][0], pair[1][0]), errs))
    return out


def reproject_poses(poses2d, camera_params, joints_vis, no_distortion=False, nviews=4,
                    return_points=False):
    """lib/multiviews/triangulate.py:169-213 -> (proj_2d like poses2d, res_vis like joints_vis)."""
    njoints = poses2d.shape[1]
    nframes = len(camera_params) // nviews
    assert np.all(joints_vis.shape == poses2d.shape[:2])
    proj_2d = np.zeros_like(poses2d)
    res_vis = np.zeros_like(joints_vis)
    pts = np.zeros((nframes, njoints, 3))
    for i in range(nframes):
        rig = _frame_rig(camera_params, i, nviews, no_distortion)
        for k in range(njoints):
            obs = _visible_obs(poses2d, joints_vis, i, k, nviews)
            if len(obs) < 2:
                continue
            X = rig.find3d(obs)
            pts[i, k] = X
            for j in range(nviews):
                proj_2d[i * nviews + j, k, :] = rig.find2d(_cam_name(j), X)
                res_vis[i * nviews + j, k] = 1
    if return_points:
        return proj_2d, res_vis, pts
    return proj_2d, res_vis
