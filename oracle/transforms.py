"""Oracle: crop affine <-> heatmap pixels (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows lib/utils/transforms.py:67-135 of the reference.  The reference calls
``cv2.getAffineTransform`` on three float32 point pairs; this restatement
builds the same three float32 point pairs and solves the 2x3 map in float64
by following OpenCV's own elimination order, so that results are bit-identical
to the reference (which uses cv2) -- pinned by tests/golden/affine.npz.
"""
import numpy as np


def _third_point(a, b):
    # lib/utils/transforms.py:123-125 -- float32 arithmetic on float32 inputs
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def _rotate_dir(pt, rot_rad):
    # lib/utils/transforms.py:128-135
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    return [pt[0] * cs - pt[1] * sn, pt[0] * sn + pt[1] * cs]


def affine_point_triples(center, scale, rot, output_size,
                         shift=np.array([0, 0], dtype=np.float32)):
    """The (src, dst) float32 [3,2] point triples of lib/utils/transforms.py:76-102."""
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])
    scale_px = scale * 200.0                      # :84 (dtype of `scale` is kept)
    src_w = scale_px[0]                           # only scale[0] is used (:85)
    dst_w, dst_h = output_size[0], output_size[1]
    src_dir = _rotate_dir([0, src_w * -0.5], np.pi * rot / 180)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_px * shift
    src[1, :] = center + src_dir + scale_px * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2, :] = _third_point(src[0, :], src[1, :])
    dst[2, :] = _third_point(dst[0, :], dst[1, :])
    return src, dst


_BACKEND = 'lu'


def set_backend(name):
    """'lu' (default, dependency-free emulation) or 'cv2' (what the reference calls).

    bench.py's CPU-baseline leg switches to 'cv2' so the timed port costs what the
    reference costs; tests/test_oracle.py checks both give identical bits.
    """
    global _BACKEND
    assert name in ('lu', 'cv2')
    _BACKEND = name


def solve_affine(frm, to):
    """2x3 float64 map with  to_k = M @ [frm_k, 1]  for the three float32 pairs.

    Restates cv2.getAffineTransform (lib/utils/transforms.py:104-107): OpenCV
    stacks the 6x6 system  [x y 1 0 0 0; 0 0 0 x y 1] m = [u; v]  and solves it by
    Gaussian elimination with partial pivoting in float64 (first largest pivot
    wins, multiplier = a_ji * (-1/a_ii), no fused multiply-add).  Following that
    operation order reproduces cv2 bit for bit (tests/golden/affine.npz: 0 of
    512 matrices differ), which the four-division closed form does not.
    """
    if _BACKEND == 'cv2':
        import cv2
        return cv2.getAffineTransform(np.float32(frm), np.float32(to))
    a = [[0.0] * 6 for _ in range(6)]
    b = [0.0] * 6
    for i in range(3):
        x, y = float(frm[i][0]), float(frm[i][1])
        a[2 * i][0], a[2 * i][1], a[2 * i][2] = x, y, 1.0
        a[2 * i + 1][3], a[2 * i + 1][4], a[2 * i + 1][5] = x, y, 1.0
        b[2 * i], b[2 * i + 1] = float(to[i][0]), float(to[i][1])
    for i in range(6):
        k = i
        for j in range(i + 1, 6):
            if abs(a[j][i]) > abs(a[k][i]):
                k = j
        if k != i:
            a[i], a[k] = a[k], a[i]
            b[i], b[k] = b[k], b[i]
        d = -1 / a[i][i]
        for j in range(i + 1, 6):
            alpha = a[j][i] * d
            for c in range(i + 1, 6):
                a[j][c] += alpha * a[i][c]
            b[j] += alpha * b[i]
    for i in range(5, -1, -1):
        s = b[i]
        for c in range(i + 1, 6):
            s -= a[i][c] * b[c]
        b[i] = s / a[i][i]
    return np.array(b).reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size,
                         shift=np.array([0, 0], dtype=np.float32), inv=0):
    """lib/utils/transforms.py:76-109."""
    src, dst = affine_point_triples(center, scale, rot, output_size, shift)
    return solve_affine(dst, src) if inv else solve_affine(src, dst)


def affine_transform(pt, t):
    """lib/utils/transforms.py:112-120 -- [pt, 1] @ t.T in float64."""
    pt = np.asarray(pt)
    if pt.ndim == 1:
        pt = pt[np.newaxis, ...]
    ones = np.ones((pt.shape[0], 1))
    return np.dot(np.concatenate((pt, ones), axis=-1), t.T).squeeze()


def transform_preds(coords, center, scale, output_size):
    """lib/utils/transforms.py:67-73 -- heatmap pixels -> image pixels."""
    out = np.zeros(coords.shape)
    t = get_affine_transform(center, scale, 0, output_size, inv=1)
    out[:, :2] = affine_transform(coords[:, :2], t)
    return out
