"""Crop affine <-> heatmap pixels, behind the names of lib/utils/transforms.py:67-135.

``get_affine_transform`` / ``transform_preds`` run the crop-affine kernel
(csrc/lift_math.cuh::crop_affine_row, bit-identical to the reference's float32 point
triples + cv2.getAffineTransform, rotations and shifts included -- pinned by
tests/golden/affine.npz).  ``crop_affine`` is the batched device form the lifting path
uses (one launch for all rows of a batch).
"""
import numpy as np
import torch

from .. import _lib, runtime as rt


def crop_affine(center, scale, output_size, inv=0, rot=None, shift=None):
    """Batched crop affine on the device: center [n,2], scale [n,2] -> CUDA tensor [n,2,3] float64.

    center / scale keep their float32 or float64 dtype: the reference's arithmetic
    depends on it (``scale * 200.0`` is rounded in the dtype of ``scale``).
    rot: degrees, a scalar or [n] (``None`` / 0 = the lifting path's unrotated crop);
    shift: the reference's ``shift`` [2] (default float32 zeros).
    """
    rt.require_device()
    c = rt.to_device_float(center).reshape(-1, 2)
    s = rt.to_device_float(scale).reshape(-1, 2)
    if c.shape != s.shape:
        raise ValueError('center %s and scale %s must both be [n, 2]' % (tuple(c.shape), tuple(s.shape)))
    n = c.shape[0]
    sincos = None
    if rot is not None and np.any(np.asarray(rot) != 0):
        # np.sin / np.cos of the float64 angle on the host, like lib/utils/transforms.py:87,129
        rad = np.pi * np.broadcast_to(np.asarray(rot, dtype=np.float64).reshape(-1), (n,)) / 180
        sincos = rt.to_device(np.stack([np.sin(rad), np.cos(rad)], axis=1))
    sh = np.zeros(2, dtype=np.float32) if shift is None else np.asarray(shift).reshape(2)
    sh_tag = _lib.F64 if sh.dtype == np.float64 or sh.dtype.kind in 'iu' else _lib.F32
    if sh.dtype not in (np.float32, np.float64):
        sh = sh.astype(np.float64)
    out = rt.empty((n, 2, 3), torch.float64)
    _lib.call('pb200_crop_affine', rt.ptr(c), rt.float_dtype_tag(c), rt.ptr(s), rt.float_dtype_tag(s),
              rt.ptr(sincos), float(sh[0]), float(sh[1]), sh_tag,
              n, int(output_size[0]), int(output_size[1]), int(bool(inv)), rt.ptr(out), rt.stream_ptr())
    return out


def get_affine_transform(center, scale, rot, output_size,
                         shift=np.array([0, 0], dtype=np.float32), inv=0):
    """lib/utils/transforms.py:76-109 -> numpy [2,3] float64."""
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])                      # transforms.py:82-83
    center = np.asarray(center)
    scale = np.asarray(scale)
    return rt.to_host(crop_affine(center.reshape(1, 2), scale.reshape(1, 2), output_size, inv,
                                  rot=rot, shift=shift))[0]


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


class _SoftArgmax2D(torch.autograd.Function):
    """softmax(beta * heatmap) expectation of (column, row): forward and backward are one HBM pass
    each in libposeb200 (csrc/softargmax.cu)."""

    @staticmethod
    def forward(ctx, heatmaps, beta):
        rt.require_device()
        hm = heatmaps.detach().contiguous()
        if hm.dtype != torch.float32 or not hm.is_cuda or hm.dim() != 4:
            raise TypeError('heatmaps must be a CUDA float32 [N, J, h, w] tensor')
        N, J, H, W = [int(v) for v in hm.shape]
        xy = torch.empty((N, J, 2), dtype=torch.float32, device=hm.device)
        stats = torch.empty((N, J, 2), dtype=torch.float32, device=hm.device)
        _lib.call('pb200_softargmax_fwd', rt.ptr(hm), N, J, H, W, float(beta), rt.ptr(xy), rt.ptr(stats),
                  rt.stream_ptr())
        ctx.save_for_backward(hm, stats, xy)
        ctx.beta = float(beta)
        return xy

    @staticmethod
    def backward(ctx, grad_xy):
        hm, stats, xy = ctx.saved_tensors
        N, J, H, W = [int(v) for v in hm.shape]
        g = grad_xy.detach().to(torch.float32).contiguous()
        grad_hm = torch.empty_like(hm)
        _lib.call('pb200_softargmax_bwd', rt.ptr(hm), rt.ptr(stats), rt.ptr(xy), rt.ptr(g), N, J, H, W,
                  ctx.beta, rt.ptr(grad_hm), rt.stream_ptr())
        return grad_hm, None


def generate_integral_preds_2d_th(heatmaps, beta=100.0):
    """lib/utils/transforms.py:149-171: differentiable soft-argmax, heatmaps [N,J,h,w] -> [N,J,2]
    (x, y) in heatmap pixels.  ``beta`` is the reference's "multiply by a factor 100"."""
    return _SoftArgmax2D.apply(heatmaps, beta)


def transform_back_th(cfg, joints_2d_list, meta):
    """lib/utils/transforms.py:174-198: heatmap pixels -> image pixels per view, differentiable.
    The per-sample inverse crop affines come from the crop-affine kernel (float64, cast to float32
    like the reference); the tiny [N,J,3] x [N,3,2] product is left to torch so autograd sees it."""
    results = []
    w, h = int(cfg.NETWORK.HEATMAP_SIZE[0]), int(cfg.NETWORK.HEATMAP_SIZE[1])
    for p, m in zip(joints_2d_list, meta):
        c = m['center'].numpy() if isinstance(m['center'], torch.Tensor) else np.asarray(m['center'])
        s = m['scale'].numpy() if isinstance(m['scale'], torch.Tensor) else np.asarray(m['scale'])
        trans = crop_affine(c, s, (w, h), inv=1).to(device=p.device, dtype=torch.float32)   # [N,2,3]
        ones = torch.ones(p.shape[0], p.shape[1], 1, device=p.device, dtype=p.dtype)
        results.append(torch.matmul(torch.cat((p, ones), dim=2), trans.transpose(2, 1)))
    return results
