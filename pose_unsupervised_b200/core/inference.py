"""Heatmap -> 2D joint decode behind the names of lib/core/inference.py:19-75.

``get_max_preds`` / ``get_final_preds`` keep the reference's signatures (numpy in,
numpy out).  ``decode_heatmaps`` is the device-resident form used where the
backbone's output is already on the GPU (lib/core/function.py:560,632-640 copies
every view to the host first; here the list of per-view CUDA tensors is decoded in
place and the rows come out view-minor).
"""
import ctypes

import numpy as np
import torch

from .. import _lib, runtime as rt
from ..utils.transforms import crop_affine


def _view_pointers(heatmaps):
    """-> (list of contiguous float32 CUDA tensors, N, J, H, W)."""
    if isinstance(heatmaps, (list, tuple)):
        views = [rt.to_device(h) for h in heatmaps]
    else:
        views = [rt.to_device(heatmaps)]
    for h in views:
        if h.dtype != torch.float32:
            raise TypeError('heatmaps must be float32 (lib/core/function.py:551 feeds float32), got %s' % h.dtype)
        if h.dim() != 4 or h.shape != views[0].shape:
            raise ValueError('heatmaps must be 4-D [n, joints, h, w] tensors of one shape')
    if len(views) > _lib.MAX_VIEWS:
        raise ValueError('at most %d view tensors' % _lib.MAX_VIEWS)
    n, j, h, w = views[0].shape
    return views, n * len(views), j, h, w


def decode_heatmaps(heatmaps, center=None, scale=None, post_process=False, return_idx=False, affine=None):
    """Decode on the device.

    heatmaps : [N,J,H,W] float32 (numpy or CUDA tensor), or a list of V per-view
               tensors [N/V,J,H,W] (rows of the result are then frame*V + view).
    center, scale : [N,2]; when given the result is in image pixels
               (get_final_preds), otherwise masked heatmap pixels (get_max_preds).
    affine   : optionally the precomputed ``crop_affine(center, scale, (W, H), inv=1)`` [N,2,3].
    Returns CUDA tensors (xy [N,J,2] float32, maxvals [N,J] float32[, idx [N,J] int32]).
    """
    rt.require_device()
    views, N, J, H, W = _view_pointers(heatmaps)
    if affine is None and center is not None:
        affine = crop_affine(center, scale, (W, H), inv=1)
    if affine is not None:
        if affine.shape[0] != N:
            raise ValueError('center/scale have %d rows, heatmaps %d' % (affine.shape[0], N))
    xy = rt.empty((N, J, 2), torch.float32)
    maxvals = rt.empty((N, J), torch.float32)
    idx = rt.empty((N, J), torch.int32) if return_idx else None
    ptrs = (ctypes.c_void_p * len(views))(*[v.data_ptr() for v in views])
    _lib.call('pb200_decode', ptrs, len(views), N, J, H, W, rt.ptr(affine), int(bool(post_process)),
              rt.ptr(xy), rt.ptr(maxvals), rt.ptr(idx), rt.stream_ptr())
    return (xy, maxvals, idx) if return_idx else (xy, maxvals)


def decode_heatmaps_flip(heatmaps, heatmaps_flipped, flip_pairs, shift_heatmap=True, center=None,
                         scale=None, post_process=False, return_idx=False):
    """Flip-test averaging + decode in one pass (lib/core/function.py:567-583, 632-640).

    heatmaps / heatmaps_flipped : the network output for the input and for the mirrored input
               (one [N,J,H,W] tensor or a list of V per-view tensors, like decode_heatmaps)
    flip_pairs : dataset.flip_pairs, e.g. [[0, 5], [1, 4], ...] (left/right joints to swap back)
    Returns (avg [N,J,H,W] float32 view-minor, xy [N,J,2], maxvals [N,J][, idx]) as CUDA tensors;
    avg equals ``(view + shift(flip_back_th(view_flipped))) * 0.5`` bit for bit.
    """
    rt.require_device()
    views, N, J, H, W = _view_pointers(heatmaps)
    fviews, Nf, Jf, Hf, Wf = _view_pointers(heatmaps_flipped)
    if (len(fviews), Nf, Jf, Hf, Wf) != (len(views), N, J, H, W):
        raise ValueError('heatmaps and heatmaps_flipped must have the same layout')
    src = np.arange(J, dtype=np.int32)
    for a, b in flip_pairs:
        src[a], src[b] = b, a
    d_src = rt.to_device(src)
    affine = None
    if center is not None:
        affine = crop_affine(center, scale, (W, H), inv=1)
        if affine.shape[0] != N:
            raise ValueError('center/scale have %d rows, heatmaps %d' % (affine.shape[0], N))
    avg = rt.empty((N, J, H, W), torch.float32)
    xy = rt.empty((N, J, 2), torch.float32)
    maxvals = rt.empty((N, J), torch.float32)
    idx = rt.empty((N, J), torch.int32) if return_idx else None
    p1 = (ctypes.c_void_p * len(views))(*[v.data_ptr() for v in views])
    p2 = (ctypes.c_void_p * len(views))(*[v.data_ptr() for v in fviews])
    _lib.call('pb200_decode_flip', p1, p2, len(views), N, J, H, W, rt.ptr(d_src), int(bool(shift_heatmap)),
              rt.ptr(affine), int(bool(post_process)), rt.ptr(avg), rt.ptr(xy), rt.ptr(maxvals), rt.ptr(idx),
              rt.stream_ptr())
    return (avg, xy, maxvals, idx) if return_idx else (avg, xy, maxvals)


def get_max_preds(batch_heatmaps):
    """lib/core/inference.py:19-47: (preds [N,J,2] float32 heatmap px, maxvals [N,J,1])."""
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    xy, maxvals = decode_heatmaps(batch_heatmaps)
    return rt.to_host(xy), rt.to_host(maxvals)[:, :, None]


def get_final_preds(config, batch_heatmaps, center, scale):
    """lib/core/inference.py:50-75: (preds [N,J,2] float32 image px, maxvals [N,J,1])."""
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    xy, maxvals = decode_heatmaps(batch_heatmaps, np.asarray(center), np.asarray(scale),
                                  post_process=bool(config.TEST.POST_PROCESS))
    return rt.to_host(xy), rt.to_host(maxvals)[:, :, None]
