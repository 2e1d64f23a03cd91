"""Oracle: the pymvg boundary, restated (TEST INFRASTRUCTURE, see oracle/__init__.py).

PARITY UNPINNED.  ``pymvg`` is the third-party package in which the reference's
triangulation arithmetic lives.  It is listed un-pinned in the reference's
requirements.txt:13 (no version, no lock file), its source is not under
/root/reference, and it is neither installed nor installable offline.  The
reference calls it at exactly four places:

* ``CameraModel.load_camera_from_M(M, name=, distortion_coefficients=)``
  lib/multiviews/triangulate.py:37-38
* ``MultiCameraSystem(cameras)``                 lib/multiviews/triangulate.py:40
* ``MultiCameraSystem.find3d(points_2d_set)``    lib/multiviews/triangulate.py:53
* ``MultiCameraSystem.find2d(name, point_3d)``   lib/multiviews/triangulate.py:147,210

This module restates pymvg's published algorithm for those calls:

``load_camera_from_M``
    RQ-decompose ``M[:, :3] = K R`` (positive diagonal, right-handed R),
    normalise by ``K[2,2]`` when it deviates from 1, recover the translation
    from the camera centre, and keep ``M = K [R | t]``.  The distortion vector
    is kept in OpenCV order ``[k1, k2, p1, p2, k3]``.  (Some pymvg releases
    lose the distortion vector in the ``K[2,2]`` re-normalisation branch; the
    oracle defines the behaviour as "distortion kept".)
``find3d``
    Hartley & Zisserman linear triangulation (H&Z 2nd ed. section 12.2): each
    observation is first undistorted with the 5-iteration fixed point of
    OpenCV's ``undistortPoints``, then contributes the rows ``x*M[2]-M[0]`` and
    ``y*M[2]-M[1]``; ``X = vt[-1,:3]/vt[-1,3]`` from ``np.linalg.svd``.  No
    Hartley normalisation -- the minimiser depends on this exact row scaling.
``find2d`` (``distorted=True`` is pymvg's default)
    ``x_c = R X + t``; pin-hole divide; OpenCV plumb-bob distortion; ``K``.

Anchors used instead of reference golden vectors (tests/test_oracle.py):
noise-free round trips recover X to <1e-6 mm, ``undistort`` agrees with
``cv2.undistortPoints`` to 1e-9 px, ``distort(undistort(x)) ~= x``.
"""
import numpy as np
import scipy.linalg


def _rq_positive(m3):
    # RQ with a positive diagonal on the upper-triangular factor
    k, r = scipy.linalg.rq(m3)
    for i in range(3):
        if k[i, i] < 0:
            k[:, i] = -k[:, i]
            r[i, :] = -r[i, :]
    return k, r


def _camera_centre(pmat):
    def minor(cols):
        return np.linalg.det(pmat[:, cols])
    x = minor([1, 2, 3])
    y = -minor([0, 2, 3])
    z = minor([0, 1, 3])
    w = -minor([0, 1, 2])
    return np.array([[x / w], [y / w], [z / w]])


class RestatedCamera(object):
    """What the reference needs from a pymvg ``CameraModel``."""

    def __init__(self, name, K, R, t, dist):
        self.name = name
        self.K = K
        self.R = R
        self.t = t.reshape(3, 1)
        self.D = np.zeros(5) if dist is None else np.asarray(dist, dtype=np.float64).reshape(5)
        self.M = K.dot(np.concatenate((R, self.t), axis=1))

    @classmethod
    def load_camera_from_M(cls, pmat, name='cam', distortion_coefficients=None,
                           _depth=0, eps=1e-15):
        pmat = np.array(pmat, dtype=np.float64)
        assert pmat.shape == (3, 4)
        K, R = _rq_positive(pmat[:, :3])
        if np.linalg.det(R) < 0:
            K, R = -K, -R
        a = K[2, 2]
        if a != 0 and abs(a - 1.0) > eps:
            if _depth > 0:
                raise ValueError('cannot scale this pmat')
            return cls.load_camera_from_M(pmat / a, name=name,
                                          distortion_coefficients=distortion_coefficients,
                                          _depth=_depth + 1, eps=max(eps, 1e-12))
        t = -R.dot(_camera_centre(pmat))
        return cls(name, K, R, t, distortion_coefficients)

    # -- lens model -------------------------------------------------------
    def undistort(self, uv):
        """[n,2] distorted pixels -> undistorted pixels (P = K)."""
        uv = np.asarray(uv, dtype=np.float64)
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        k1, k2, p1, p2, k3 = self.D
        xd = (uv[:, 0] - cx) / fx
        yd = (uv[:, 1] - cy) / fy
        x, y = xd.copy(), yd.copy()
        for _ in range(5):
            r2 = x * x + y * y
            icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
            dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
            dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
            x = (xd - dx) * icdist
            y = (yd - dy) * icdist
        return np.stack([x * fx + cx, y * fy + cy], axis=1)

    def distort(self, uv):
        """[n,2] ideal pixels -> distorted pixels (plumb-bob)."""
        uv = np.asarray(uv, dtype=np.float64)
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        k1, k2, p1, p2, k3 = self.D
        x = (uv[:, 0] - cx) / fx
        y = (uv[:, 1] - cy) / fy
        r2 = x * x + y * y
        r4 = r2 * r2
        r6 = r4 * r2
        a1 = 2 * x * y
        barrel = 1 + k1 * r2 + k2 * r4 + k3 * r6
        xpp = x * barrel + p1 * a1 + p2 * (r2 + 2 * (x * x))
        ypp = y * barrel + p1 * (r2 + 2 * (y * y)) + p2 * a1
        return np.stack([xpp * fx + cx, ypp * fy + cy], axis=1)

    def project_3d_to_pixel(self, pts3d, distorted=True):
        pts3d = np.asarray(pts3d, dtype=np.float64).reshape(-1, 3)
        cc = self.R.dot(pts3d.T) + self.t          # [3,n]
        hom = self.K.dot(cc)
        uv = (hom[:2] / hom[2]).T
        return self.distort(uv) if distorted else uv


class RestatedMultiCameraSystem(object):
    """What the reference needs from a pymvg ``MultiCameraSystem``."""

    def __init__(self, cameras):
        self._cams = {}
        for c in cameras:
            assert c.name not in self._cams, 'camera names must be unique'
            self._cams[c.name] = c

    def find3d(self, pts, undistort=True):
        rows = []
        for name, xy in pts:
            cam = self._cams[name]
            xy = np.asarray(xy, dtype=np.float64).reshape(1, 2)
            if undistort:
                xy = cam.undistort(xy)
            x, y = xy[0]
            rows.append(x * cam.M[2] - cam.M[0])
            rows.append(y * cam.M[2] - cam.M[1])
        _, _, vt = np.linalg.svd(np.array(rows))
        return vt[-1, 0:3] / vt[-1, 3]

    def find2d(self, name, xyz, distorted=True):
        return self._cams[name].project_3d_to_pixel(xyz, distorted=distorted)[0]
