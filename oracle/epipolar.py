"""Oracle: epipolar-consistency residual (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows run/test/test_fund_mtx.py:56-69 (float64 evaluation script) and
lib/core/loss.py:101-133 (``FundamentalLoss``; same formula, torch float32):
for every ordered view pair (a, b) of ``itertools.permutations(range(V), 2)``
and every joint,  r = | [x_b, 1] F_(subject,a,b) . [x_a, 1] |  -- the algebraic
residual, not Sampson.  The reference reads F from a pickle that is not
shipped (``cv2.findFundamentalMat`` LMedS per subject and pair,
run/test/generate_fundamental_matirx.py:45-57); here F is exact from the
cameras so that ``x_b^T F x_a = 0`` for undistorted projections.
"""
import itertools

import numpy as np


def view_pairs(nviews):
    return list(itertools.permutations(range(nviews), 2))


def fundamental_from_cameras(cam_a, cam_b):
    """F with x_b^T F x_a = 0 for pin-hole (undistorted) projections of one point."""
    def K(cam):
        return np.array([[float(np.ravel(cam['fx'])[0]), 0, float(np.ravel(cam['cx'])[0])],
                         [0, float(np.ravel(cam['fy'])[0]), float(np.ravel(cam['cy'])[0])],
                         [0, 0, 1.0]])
    Ra, Rb = cam_a['R'], cam_b['R']
    Ca, Cb = np.reshape(cam_a['T'], 3), np.reshape(cam_b['T'], 3)
    R = Rb.dot(Ra.T)                              # x_b = R x_a + t
    t = Rb.dot(Ca - Cb)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    F = np.linalg.inv(K(cam_b)).T.dot(tx).dot(R).dot(np.linalg.inv(K(cam_a)))
    return F / np.linalg.norm(F)


def fundamental_table(cams_by_subject):
    """{(subject, a, b): F [3,3]} like the reference's fundamental_matrix.pkl."""
    out = {}
    for subj, cams in cams_by_subject.items():
        for a, b in view_pairs(len(cams)):
            out[(subj, a, b)] = fundamental_from_cameras(cams[a], cams[b])
    return out


def epipolar_residuals(pred2d, subjects, fmat, nviews=4):
    """run/test/test_fund_mtx.py:56-69 -> |residual| [B, V(V-1), J] float64.

    pred2d [B*V, J, 2] view-minor; subjects [B]; fmat {(subj,a,b): F}.
    """
    njoints = pred2d.shape[1]
    batches = np.reshape(pred2d, (len(pred2d) // nviews, nviews, njoints, 2))
    res = []
    for subj, batch in zip(subjects, batches):
        row = []
        for a, b in view_pairs(nviews):
            pa = np.concatenate((batch[a], np.ones((njoints, 1))), axis=1)
            pb = np.concatenate((batch[b], np.ones((njoints, 1))), axis=1)
            row.append(np.sum((pb @ fmat[(subj, a, b)]) * pa, axis=1))
        res.append(row)
    return np.abs(np.array(res))


def fundamental_loss(joints_2d_list, target_weight, subjects, fmat, use_target_weight):
    """lib/core/loss.py:101-133 in numpy float64.

    joints_2d_list: V arrays [K,J,2]; target_weight: V arrays [K,J,1].
    """
    nviews = len(joints_2d_list)
    k, j = joints_2d_list[0].shape[:2]
    homo = [np.concatenate((p, np.ones((k, j, 1))), axis=2) for p in joints_2d_list]
    pairs = view_pairs(nviews)
    loss = 0.0
    for i, subj in enumerate(subjects):
        for a, b in pairs:
            r = np.abs(np.sum((homo[b][i] @ fmat[(subj, a, b)]) * homo[a][i], axis=1))
            if use_target_weight:
                r = r * np.squeeze(target_weight[b][i] * target_weight[a][i])
            loss += r.sum()
    return loss / (k * len(pairs) * j)
