"""Oracle: H36M pin-hole camera with radial/tangential distortion.

TEST INFRASTRUCTURE, see oracle/__init__.py.  Restates the arithmetic of
lib/multiviews/cameras.py:12-82 (same operation order, so results are
bit-identical to the reference -- pinned by tests/golden/cameras.npz).

Camera dicts carry ``R [3,3]``, ``T [3,1]`` (camera centre in world
coordinates), ``fx fy cx cy`` (shape-(1,) arrays, the H36M cameras.h5 style),
``k [3,1]`` radial and ``p [2,1]`` tangential coefficients.
"""
import numpy as np


def camera_fields(cam, avg_f=True):
    """(R, T, f, c, k, p) as lib/multiviews/cameras.py:12-22 unfolds them."""
    fx, fy = cam['fx'], cam['fy']
    f = 0.5 * (fx + fy) if avg_f else np.array([fx, fy])
    c = np.array([cam['cx'], cam['cy']])
    return cam['R'], cam['T'], f, c, cam['k'], cam['p']


def project_h36m(pts, R, T, f, c, k, p):
    """World points [n,3] -> pixels [n,2]; lib/multiviews/cameras.py:25-49.

    The tangential term is the H36M form: the multiplicative part is
    ``p0*y1 + p1*y0`` and the additive part ``[p1, p0] * r2``.  This is NOT the
    OpenCV plumb-bob model the pymvg reprojection uses
    (oracle/pymvg_restated.py); the two differ by up to ~0.3 px.
    """
    k = np.asarray(k, dtype=np.float64).reshape(3)
    p = np.asarray(p, dtype=np.float64).reshape(2)
    cam_xyz = R.dot(pts.T - T)                     # [3,n]
    u = cam_xyz[0] / cam_xyz[2]
    v = cam_xyz[1] / cam_xyz[2]
    r2 = u ** 2 + v ** 2
    poly = k[0] * r2 + k[1] * r2 ** 2 + k[2] * r2 ** 3
    gain = (1 + poly) + (p[0] * v + p[1] * u)
    ud = u * gain + p[1] * r2
    vd = v * gain + p[0] * r2
    f = np.asarray(f, dtype=np.float64).reshape(-1)
    c = np.asarray(c, dtype=np.float64).reshape(2)
    fx, fy = (f[0], f[0]) if f.size == 1 else (f[0], f[1])
    return np.stack([fx * ud + c[0], fy * vd + c[1]], axis=1)


def project_pose(pts, cam):
    """lib/multiviews/cameras.py:52-54: averaged focal length."""
    return project_h36m(pts, *camera_fields(cam, avg_f=True))


def world_to_camera_frame(pts, R, T):
    """lib/multiviews/cameras.py:57-68:  R (x - T)."""
    return R.dot(pts.T - T).T


def camera_to_world_frame(pts, R, T):
    """lib/multiviews/cameras.py:71-82:  R^T x + T."""
    return (R.T.dot(pts.T) + T).T
