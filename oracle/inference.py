"""Oracle: heatmap -> 2D joint decode (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows lib/core/inference.py:19-75 of the reference.  Two forms are kept:

* ``*_loops``  -- the reference's own control flow (numpy argmax/amax over the
  flattened map, then the per-(sample, joint) Python refinement loop and the
  per-sample inverse affine).  This is what the reference costs on a CPU and is
  the form timed by bench.py's cpu_baseline.
* vectorised   -- same arithmetic with the Python loops replaced by array
  expressions, so that the checker finishes in seconds at parity-test sizes.
  tests/test_oracle.py checks the two forms are identical.
"""
import math

import numpy as np

from .transforms import transform_preds, get_affine_transform


def get_max_preds(batch_heatmaps):
    """lib/core/inference.py:19-47.  Returns (preds [N,J,2] f32, maxvals [N,J,1])."""
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    n, j, _, w = batch_heatmaps.shape
    flat = batch_heatmaps.reshape((n, j, -1))
    idx = np.argmax(flat, 2)                      # first maximal index; NaN counts as max
    maxvals = np.amax(flat, 2).reshape((n, j, 1))
    idx_f = idx.astype(np.float32)                # exact below 2**24
    preds = np.empty((n, j, 2), dtype=np.float32)
    preds[:, :, 0] = idx_f % w
    preds[:, :, 1] = np.floor(idx_f / w)
    preds *= np.greater(maxvals, 0.0).astype(np.float32)
    return preds, maxvals


def flat_argmax(batch_heatmaps):
    """The integer index behind get_max_preds (bit-exact target for the CUDA kernel)."""
    n, j = batch_heatmaps.shape[:2]
    return np.argmax(batch_heatmaps.reshape((n, j, -1)), 2)


def quarter_pixel_loops(coords, batch_heatmaps):
    """lib/core/inference.py:57-66, in place, the reference's double loop."""
    h, w = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    for n in range(coords.shape[0]):
        for p in range(coords.shape[1]):
            hm = batch_heatmaps[n][p]
            px = int(math.floor(coords[n][p][0] + 0.5))
            py = int(math.floor(coords[n][p][1] + 0.5))
            if 1 < px < w - 1 and 1 < py < h - 1:
                diff = np.array([hm[py][px + 1] - hm[py][px - 1],
                                 hm[py + 1][px] - hm[py - 1][px]])
                coords[n][p] += np.sign(diff) * .25
    return coords


def quarter_pixel(coords, batch_heatmaps):
    """Vectorised form of :func:`quarter_pixel_loops` (same arithmetic)."""
    n, j, h, w = batch_heatmaps.shape
    px = np.floor(coords[:, :, 0] + np.float32(0.5)).astype(np.int64)
    py = np.floor(coords[:, :, 1] + np.float32(0.5)).astype(np.int64)
    ok = (px > 1) & (px < w - 1) & (py > 1) & (py < h - 1)
    pxc = np.clip(px, 1, w - 2)
    pyc = np.clip(py, 1, h - 2)
    ni, ji = np.meshgrid(np.arange(n), np.arange(j), indexing='ij')
    dx = batch_heatmaps[ni, ji, pyc, pxc + 1] - batch_heatmaps[ni, ji, pyc, pxc - 1]
    dy = batch_heatmaps[ni, ji, pyc + 1, pxc] - batch_heatmaps[ni, ji, pyc - 1, pxc]
    step = np.stack([np.sign(dx), np.sign(dy)], axis=2) * .25
    coords += np.where(ok[:, :, None], step, 0.0).astype(coords.dtype)
    return coords


def get_final_preds_loops(post_process, batch_heatmaps, center, scale):
    """lib/core/inference.py:50-75 with the reference's loops (CPU-baseline form)."""
    coords, maxvals = get_max_preds(batch_heatmaps)
    h, w = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    if post_process:
        quarter_pixel_loops(coords, batch_heatmaps)
    preds = coords.copy()
    for i in range(coords.shape[0]):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [w, h])
    return preds, maxvals


def get_final_preds(post_process, batch_heatmaps, center, scale):
    """Checker form: vectorised refinement, per-sample affine.

    ``post_process`` plays the role of ``config.TEST.POST_PROCESS``.
    Returns (preds [N,J,2] f32 in image pixels, maxvals [N,J,1]).
    """
    coords, maxvals = get_max_preds(batch_heatmaps)
    h, w = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    if post_process:
        quarter_pixel(coords, batch_heatmaps)
    preds = coords.copy()
    for i in range(coords.shape[0]):
        t = get_affine_transform(center[i], scale[i], 0, [w, h], inv=1)
        xy1 = np.concatenate((coords[i, :, :2], np.ones((coords.shape[1], 1))), axis=-1)
        preds[i] = np.dot(xy1, t.T)
    return preds, maxvals
