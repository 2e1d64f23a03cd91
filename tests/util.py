"""Shared helpers for the tests (golden loaders, camera packing, config stubs)."""
import os
import types

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def unpack_cam(v):
    """21-float golden record -> camera dict."""
    z = np.zeros(1)
    return {'R': v[:9].reshape(3, 3).copy(), 'T': v[9:12].reshape(3, 1).copy(),
            'fx': z + v[12], 'fy': z + v[13], 'cx': z + v[14], 'cy': z + v[15],
            'k': v[16:19].reshape(3, 1).copy(), 'p': v[19:21].reshape(2, 1).copy()}


def rpsm_config(first=16, recur=2, depth=10, grid=2000, tol=150, img=256, hm=64):
    return types.SimpleNamespace(
        NETWORK=types.SimpleNamespace(IMAGE_SIZE=np.array([img, img]), HEATMAP_SIZE=np.array([hm, hm])),
        PICT_STRUCT=types.SimpleNamespace(FIRST_NBINS=first, RECUR_NBINS=recur, RECUR_DEPTH=depth,
                                          GRID_SIZE=grid, LIMB_LENGTH_TOLERANCE=tol))


def pseudo_config(reproj_thre=10.0, num_inliers=3, no_distortion=False):
    return types.SimpleNamespace(
        DATASET=types.SimpleNamespace(NO_DISTORTION=no_distortion),
        PSEUDO_LABEL=types.SimpleNamespace(REPROJ_THRE=reproj_thre, NUM_INLIERS=num_inliers))


def decode_config(post_process):
    return types.SimpleNamespace(TEST=types.SimpleNamespace(POST_PROCESS=post_process))


def rpsm_golden_frame(r, f):
    """Inputs of golden RPSM frame f: (heatmaps, cams, boxes, root, limb dict, edges)."""
    edges = [tuple(int(x) for x in e) for e in r['edges']]
    hm = r['f%d_hm_q12' % f].astype(np.float32) / np.float32(4096)
    cams = [unpack_cam(v) for v in r['f%d_cams' % f]]
    boxes = [{'center': a, 'scale': b} for a, b in zip(r['f%d_box_center' % f], r['f%d_box_scale' % f])]
    limb = {e: float(l) for e, l in zip(edges, r['f%d_limb' % f])}
    return hm, cams, boxes, r['f%d_root' % f], limb, edges


def ulp_diff_f32(a, b):
    """Elementwise distance in float32 ulps (NaN == NaN)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2 ** 31) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2 ** 31) - ib, ib)
    d = np.abs(ia - ib)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)
