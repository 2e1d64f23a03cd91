"""Crop affine <-> heatmap pixels, behind the names of lib/utils/transforms.py:67-135.

``get_affine_transform`` / ``transform_preds`` run the crop-affine kernel
(csrc/lift_math.cuh::crop_affine_row, bit-identical to cv2.getAffineTransform on
the reference's float32 point triples).  Only the configuration the lifting path
uses is implemented: ``rot == 0`` and ``shift == 0`` (lib/core/inference.py:70-73,
lib/multiviews/pictorial.py:169-170); data-augmentation rotations stay with the
reference's own module.
"""
import numpy as np
import torch

from .. import _lib, runtime as rt


def crop_affine(center, scale, output_size, inv=0):
    """Batched crop affine on the device: center [n,2], scale [n,2] -> CUDA tensor [n,2,3] float64.

    center / scale keep their float32 or float64 dtype: the reference's arithmetic
    depends on it (``scale * 200.0`` is rounded in the dtype of ``scale``).
    """
    rt.require_device()
    c = rt.to_device_float(center).reshape(-1, 2)
    s = rt.to_device_float(scale).reshape(-1, 2)
    if c.shape != s.shape:
        raise ValueError('center %s and scale %s must both be [n, 2]' % (tuple(c.shape), tuple(s.shape)))
    n = c.shape[0]
    out = rt.empty((n, 2, 3), torch.float64)
    _lib.call('pb200_crop_affine', rt.ptr(c), rt.float_dtype_tag(c), rt.ptr(s), rt.float_dtype_tag(s),
              n, int(output_size[0]), int(output_size[1]), int(bool(inv)), rt.ptr(out), rt.stream_ptr())
    return out


def get_affine_transform(center, scale, rot, output_size,
                         shift=np.array([0, 0], dtype=np.float32), inv=0):
    """lib/utils/transforms.py:76-109 -> numpy [2,3] float64 (rot = 0, shift = 0 only)."""
    if rot != 0 or np.any(np.asarray(shift) != 0):
        raise NotImplementedError('only rot=0, shift=0 is on the lifting path (see module docstring)')
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])                      # transforms.py:82-83
    center = np.asarray(center)
    scale = np.asarray(scale)
    return rt.to_host(crop_affine(center.reshape(1, 2), scale.reshape(1, 2), output_size, inv))[0]


def affine_transform(pt, t):
    """lib/utils/transforms.py:112-120: [pt, 1] @ t.T (host helper, not on the hot path)."""
    pt = np.asarray(pt)
    if pt.ndim == 1:
        pt = pt[np.newaxis, ...]
    return np.dot(np.concatenate((pt, np.ones((pt.shape[0], 1))), axis=-1), t.T).squeeze()


def transform_preds(coords, center, scale, output_size):
    """lib/utils/transforms.py:67-73: heatmap pixels [J,2+] -> image pixels, float64."""
    coords = np.asarray(coords)
    t = crop_affine(np.asarray(center).reshape(1, 2), np.asarray(scale).reshape(1, 2), output_size, inv=1)
    xy = rt.to_device_float(coords[:, :2]).reshape(1, -1, 2)
    out = rt.empty(xy.shape, torch.float64)
    _lib.call('pb200_transform_preds', rt.ptr(xy), rt.float_dtype_tag(xy), rt.ptr(t), 1, xy.shape[1],
              rt.ptr(out), rt.stream_ptr())
    target = np.zeros(coords.shape)
    target[:, :2] = rt.to_host(out)[0]
    return target
