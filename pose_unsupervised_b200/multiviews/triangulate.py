"""Multi-view triangulation, RANSAC view selection and reprojection behind the names of
lib/multiviews/triangulate.py:57-213, plus the fused heatmap -> 3D pass.

The reference builds a pymvg camera rig per frame and loops frames x joints in
Python; here one CUDA thread owns one (frame, joint) (csrc/geometry.cu), or, for
``lift_heatmaps``, the decode warps lift each frame as soon as its last heatmap has
been decoded (csrc/lift_fused.cu).  ``nviews`` is a keyword (the reference
hard-codes 4 at triangulate.py:70,114,183).
"""
import ctypes

import numpy as np
import torch

from .. import _lib, runtime as rt
from ..core.inference import _view_pointers
from ..utils.transforms import crop_affine
from .cameras import CameraTable


def _prep(camera_params, poses2d, joints_vis, nviews):
    rt.require_device()
    table = CameraTable.from_cameras(camera_params)
    xy = rt.to_device_float(poses2d)
    if xy.dim() != 3 or xy.shape[2] != 2:
        raise ValueError('poses2d must be [N, k, 2]')
    N, J = int(xy.shape[0]), int(xy.shape[1])
    B = len(table) // nviews                                   # triangulate.py:72
    if B * nviews > N:
        raise ValueError('%d cameras for %d rows of poses2d' % (len(table), N))
    vis = None
    if joints_vis is not None:
        jv = joints_vis if isinstance(joints_vis, torch.Tensor) else np.asarray(joints_vis)
        assert tuple(jv.shape) == (N, J), 'joints_vis must be [N, k]'   # triangulate.py:74
        vis = (rt.to_device(jv) != 0).to(torch.uint8)            # python truthiness of the reference
    return table, xy, vis, B, J


def triangulate_poses(camera_params, poses2d, joints_vis=None, no_distortion=False, nviews=4):
    """lib/multiviews/triangulate.py:57-99 -> poses3d [N/nviews, k, 3] float64."""
    table, xy, vis, B, J = _prep(camera_params, poses2d, joints_vis, nviews)
    out = rt.empty((B, J, 3), torch.float64)
    _lib.call('pb200_triangulate', rt.ptr(table.pack), rt.ptr(table.index), rt.ptr(xy),
              rt.float_dtype_tag(xy), rt.ptr(vis), B, nviews, J, int(bool(no_distortion)),
              rt.ptr(out), rt.stream_ptr())
    return out if rt.is_device_tensor(poses2d) else rt.to_host(out)


def ransac(poses2d, camera_params, joints_vis, config, nviews=4):
    """lib/multiviews/triangulate.py:102-166 -> res_vis like joints_vis."""
    table, xy, vis, B, J = _prep(camera_params, poses2d, joints_vis, nviews)
    out = rt.zeros((xy.shape[0], J), torch.uint8)
    _lib.call('pb200_ransac', rt.ptr(table.pack), rt.ptr(table.index), rt.ptr(xy),
              rt.float_dtype_tag(xy), rt.ptr(vis), B, nviews, J,
              int(bool(config.DATASET.NO_DISTORTION)), float(config.PSEUDO_LABEL.REPROJ_THRE),
              int(config.PSEUDO_LABEL.NUM_INLIERS), rt.ptr(out), rt.stream_ptr())
    if rt.is_device_tensor(joints_vis):
        return out.to(joints_vis.dtype)
    return rt.to_host(out).astype(np.asarray(joints_vis).dtype)   # np.zeros_like(joints_vis)


def reproject_poses(poses2d, camera_params, joints_vis, no_distortion=False, nviews=4,
                    return_points=False):
    """lib/multiviews/triangulate.py:169-213 -> (proj_2d like poses2d, res_vis like joints_vis)."""
    table, xy, vis, B, J = _prep(camera_params, poses2d, joints_vis, nviews)
    assert vis is not None, 'joints_vis is required'             # triangulate.py:186
    N = int(xy.shape[0])
    proj = rt.zeros((N, J, 2), torch.float64)
    res_vis = rt.zeros((N, J), torch.uint8)
    pts = rt.empty((B, J, 3), torch.float64) if return_points else None
    _lib.call('pb200_reproject', rt.ptr(table.pack), rt.ptr(table.index), rt.ptr(xy),
              rt.float_dtype_tag(xy), rt.ptr(vis), B, nviews, J, int(bool(no_distortion)),
              rt.ptr(proj), rt.ptr(res_vis), rt.ptr(pts), None, rt.stream_ptr())
    if rt.is_device_tensor(poses2d):
        res = (proj.to(xy.dtype), res_vis.to(joints_vis.dtype) if rt.is_device_tensor(joints_vis) else res_vis)
        return res + (pts,) if return_points else res
    res = (rt.to_host(proj).astype(np.asarray(poses2d).dtype),
           rt.to_host(res_vis).astype(np.asarray(joints_vis).dtype))
    return res + (rt.to_host(pts),) if return_points else res


class LiftResult(object):
    """Outputs of :func:`lift_heatmaps` (CUDA tensors; ``.numpy()`` copies them to the host)."""

    def __init__(self, xy, maxvals, idx, poses3d, reproj_err, proj2d, epipolar=None):
        self.xy, self.maxvals, self.idx = xy, maxvals, idx
        self.poses3d, self.reproj_err, self.proj2d = poses3d, reproj_err, proj2d
        self.epipolar = epipolar          # [B, V(V-1), J] float64 or None

    def numpy(self):
        f = lambda t: None if t is None else t.cpu().numpy()
        return LiftResult(f(self.xy), f(self.maxvals), f(self.idx), f(self.poses3d),
                          f(self.reproj_err), f(self.proj2d), f(self.epipolar))


def lift_heatmaps(heatmaps, center, scale, camera_params, nviews=4, post_process=True,
                  no_distortion=False, conf_thre=None, return_idx=False, return_proj=False,
                  affine=None, out_poses3d=None, fundamental=None, subjects=None):
    """Heatmaps -> 2D joints -> 3D poses -> reprojection error in one pass over HBM.

    Equivalent to ``get_final_preds`` (lib/core/inference.py:50-75) on every row followed by
    ``reproject_poses`` (lib/multiviews/triangulate.py:169-213) on the decoded coordinates with
    ``joints_vis = maxvals > conf_thre`` (run/test/test_pseudo_label.py:194; all visible when
    ``conf_thre`` is None).  heatmaps: [B*V,J,H,W] float32 view-minor, or a list of V
    per-view tensors [B,J,H,W].  ``affine`` may carry the [N,2,3] result of
    ``crop_affine(center, scale, (W, H), inv=1)`` when the caller already has it;
    ``out_poses3d`` a preallocated CUDA float64 [B,J,3] tensor to write the poses into (e.g. the
    send buffer of ``parallel.PoseExchange``).  With ``fundamental`` (a ``core.loss.FundamentalTable``)
    and ``subjects`` [B], the algebraic epipolar residuals of the decoded coordinates
    (run/test/test_fund_mtx.py:56-69) are produced in the same pass as ``result.epipolar``.
    """
    rt.require_device()
    views, N, J, H, W = _view_pointers(heatmaps)
    if N % nviews != 0:
        raise ValueError('%d rows are not a multiple of nviews=%d' % (N, nviews))
    if len(views) not in (1, nviews):
        raise ValueError('pass one [B*V,...] tensor or exactly nviews per-view tensors')
    B = N // nviews
    table = CameraTable.from_cameras(camera_params)
    if len(table) < N:
        raise ValueError('%d cameras for %d rows' % (len(table), N))
    if affine is None:
        affine = crop_affine(center, scale, (W, H), inv=1)
    if affine.shape[0] != N:
        raise ValueError('center/scale have %d rows, heatmaps %d' % (affine.shape[0], N))
    xy = rt.empty((N, J, 2), torch.float32)
    maxvals = rt.empty((N, J), torch.float32)
    idx = rt.empty((N, J), torch.int32) if return_idx else None
    if out_poses3d is None:
        poses3d = rt.empty((B, J, 3), torch.float64)
    else:
        poses3d = out_poses3d
        if poses3d.dtype != torch.float64 or tuple(poses3d.shape) != (B, J, 3) or \
                not poses3d.is_cuda or not poses3d.is_contiguous():
            raise ValueError('out_poses3d must be a contiguous CUDA float64 [%d, %d, 3] tensor' % (B, J))
    err = rt.empty((N, J), torch.float32)
    proj = rt.empty((N, J, 2), torch.float64) if return_proj else None
    fmat = slots = resid = None
    if fundamental is not None:
        if subjects is None:
            raise ValueError('subjects [B] are needed with a fundamental table')
        fmat, slots = fundamental.fmat, fundamental.slots(subjects)
        if int(slots.shape[0]) != B or fundamental.nviews != nviews:
            raise ValueError('fundamental table / subjects do not match the batch')
        resid = rt.empty((B, nviews * (nviews - 1), J), torch.float64)
    ws = rt.workspace('lift', _lib.load().pb200_lift_workspace_bytes(B, nviews, J))
    ptrs = (ctypes.c_void_p * len(views))(*[v.data_ptr() for v in views])
    _lib.call('pb200_lift_fused', ptrs, len(views), B, nviews, J, H, W, rt.ptr(affine),
              int(bool(post_process)), rt.ptr(table.pack), rt.ptr(table.index),
              int(bool(no_distortion)), int(conf_thre is not None),
              float(0.0 if conf_thre is None else conf_thre),
              rt.ptr(xy), rt.ptr(maxvals), rt.ptr(idx), rt.ptr(poses3d), rt.ptr(err), rt.ptr(proj),
              rt.ptr(fmat), rt.ptr(slots), rt.ptr(resid), rt.ptr(ws), rt.stream_ptr())
    return LiftResult(xy, maxvals, idx, poses3d, err, proj, resid)


def mpjpe_stats(pred3d, gt3d, out=None):
    """Partial sums of run/test/test_triangulate.py:98-101 on the device.

    Returns a CUDA float64 tensor [sum, sum of squares, max, count] over the [B,J] joint
    errors |pred - gt|; accumulate several shards by passing ``out`` again.  This is the
    all-reduce payload of parallel.py.
    """
    rt.require_device()
    p = rt.to_device(pred3d, torch.float64)
    g = rt.to_device(gt3d, torch.float64)
    if p.shape != g.shape or p.dim() != 3 or p.shape[2] != 3:
        raise ValueError('pred3d and gt3d must both be [B, J, 3]')
    if out is None:
        out = rt.zeros((4,), torch.float64)
    _lib.call('pb200_mpjpe_stats', rt.ptr(p), rt.ptr(g), int(p.shape[0]), int(p.shape[1]),
              rt.ptr(out), rt.stream_ptr())
    return out


def mpjpe_summary(stats):
    """[sum, sumsq, max, count] -> dict(mean, std, max) as printed by test_triangulate.py:99-101."""
    s, s2, mx, n = [float(v) for v in (stats.cpu() if isinstance(stats, torch.Tensor) else stats)]
    mean = s / n
    return {'mean': mean, 'std': float(np.sqrt(max(s2 / n - mean * mean, 0.0))), 'max': mx, 'count': n}
