"""Multi-view triangulation, RANSAC view selection and reprojection behind the names of
lib/multiviews/triangulate.py:57-213, plus the fused heatmap -> 3D pass.

The reference builds a pymvg camera rig per frame and loops frames x joints in
Python; here one CUDA thread owns one (frame, joint) (csrc/geometry.cu); ``lift_heatmaps``
runs the decode and that per-joint lift back to back on the device (csrc/lift_fused.cu).  ``nviews`` is a keyword (the reference
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


def _lift_launch(views, B, nviews, J, H, W, affine, post_process, pack, index, no_distortion, conf_thre,
                 xy, maxvals, idx, poses3d, err, proj, fmat, slots, resid, after_decode=None):
    """One pb200_lift_fused call on device tensors (outputs preallocated by the caller); with
    ``after_decode`` the two launches are issued separately and the callable runs in between."""
    ptrs = (ctypes.c_void_p * len(views))(*[v.data_ptr() for v in views])
    if after_decode is not None:
        _lib.call('pb200_decode', ptrs, len(views), B * nviews, J, H, W, rt.ptr(affine), int(bool(post_process)),
                  rt.ptr(xy), rt.ptr(maxvals), rt.ptr(idx), rt.stream_ptr())
        after_decode()
        _lib.call('pb200_lift_decoded', rt.ptr(pack), rt.ptr(index), rt.ptr(xy), rt.ptr(maxvals),
                  int(conf_thre is not None), float(0.0 if conf_thre is None else conf_thre), B, nviews, J,
                  int(bool(no_distortion)), rt.ptr(poses3d), rt.ptr(err), rt.ptr(proj), rt.ptr(fmat),
                  rt.ptr(slots), rt.ptr(resid), rt.stream_ptr())
        return
    _lib.call('pb200_lift_fused', ptrs, len(views), B, nviews, J, H, W, rt.ptr(affine),
              int(bool(post_process)), rt.ptr(pack), rt.ptr(index),
              int(bool(no_distortion)), int(conf_thre is not None),
              float(0.0 if conf_thre is None else conf_thre),
              rt.ptr(xy), rt.ptr(maxvals), rt.ptr(idx), rt.ptr(poses3d), rt.ptr(err), rt.ptr(proj),
              rt.ptr(fmat), rt.ptr(slots), rt.ptr(resid), rt.stream_ptr())


# Host heatmaps in pageable memory are staged through two pinned buffers in chunks of this many
# bytes: the host copy of chunk k+1 runs while chunk k crosses PCIe and chunk k-1 is decoded.
STAGE_BYTES = 256 << 20
PAGEABLE_MIN_BYTES = 64 << 20     # smaller arrays go over in one plain copy
_stage = {}


def _staging(nbytes):
    key = torch.cuda.current_device()
    st = _stage.get(key)
    if st is None or st['pinned'][0].numel() < nbytes:
        st = {'pinned': [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)],
              'device': [torch.empty(nbytes, dtype=torch.uint8, device=rt.device()) for _ in range(2)],
              'copy_stream': torch.cuda.Stream(),
              'h2d_done': [torch.cuda.Event() for _ in range(2)],
              'lift_done': [torch.cuda.Event() for _ in range(2)],
              'pool': None}
        _stage[key] = st
    return st


def _host_copy(dst, src, pool, nthreads):
    """memcpy of a large numpy block with several threads (numpy releases the GIL while copying)."""
    n = src.shape[0]
    if pool is None or n < 2 * nthreads:
        np.copyto(dst, src)
        return
    step = (n + nthreads - 1) // nthreads
    futs = [pool.submit(np.copyto, dst[i:i + step], src[i:i + step]) for i in range(0, n, step)]
    for f in futs:
        f.result()


def _lift_from_pageable(hm, B, nviews, J, H, W, affine, post_process, table, no_distortion, conf_thre,
                        outs, fmat, slots):
    """Chunked, double-buffered host -> device pipeline for pageable numpy heatmaps."""
    import concurrent.futures
    import os
    xy, maxvals, idx, poses3d, err, proj, resid = outs
    frame_bytes = nviews * J * H * W * 4
    fpc = max(1, min(B, STAGE_BYTES // frame_bytes))             # frames per chunk
    st = _staging(fpc * frame_bytes)
    if st['pool'] is None:
        st['pool'] = concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
    nthreads = st['pool']._max_workers
    main = torch.cuda.current_stream()
    flat = hm.reshape(B, -1)
    for c, lo in enumerate(range(0, B, fpc)):
        hi = min(B, lo + fpc)
        s = c & 1
        n = hi - lo
        pinned = st['pinned'][s][:n * frame_bytes].view(torch.float32).view(n, -1)
        dev = st['device'][s][:n * frame_bytes].view(torch.float32)
        st['h2d_done'][s].synchronize()                           # pinned buffer s is free again
        _host_copy(pinned.numpy(), flat[lo:hi], st['pool'], nthreads)
        with torch.cuda.stream(st['copy_stream']):
            if c >= 2:
                st['copy_stream'].wait_event(st['lift_done'][s])  # device buffer s has been decoded
            else:
                st['copy_stream'].wait_stream(main)
            dev.copy_(pinned.view(-1), non_blocking=True)
            st['h2d_done'][s].record(st['copy_stream'])
        main.wait_event(st['h2d_done'][s])
        r0, r1 = lo * nviews, hi * nviews
        _lift_launch([dev.view(n * nviews, J, H, W)], n, nviews, J, H, W, affine[r0:r1], post_process,
                     table.pack, table.index[r0:r1], no_distortion, conf_thre,
                     xy[r0:r1], maxvals[r0:r1], None if idx is None else idx[r0:r1], poses3d[lo:hi],
                     err[r0:r1], None if proj is None else proj[r0:r1], fmat,
                     None if slots is None else slots[lo:hi], None if resid is None else resid[lo:hi])
        st['lift_done'][s].record(main)


def lift_heatmaps(heatmaps, center, scale, camera_params, nviews=4, post_process=True,
                  no_distortion=False, conf_thre=None, return_idx=False, return_proj=False,
                  affine=None, out_poses3d=None, fundamental=None, subjects=None, after_decode=None):
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

    ``after_decode``: a callable run between the decode launch and the lift launch (device-resident
    input only); ``parallel.PoseExchange.pipelined_step`` uses it to start the previous step's
    all-gather underneath the latency-bound lift kernel.

    A numpy array in pageable host memory (what ``validate()`` hands over after ``.cpu().numpy()``,
    lib/core/function.py:633-640) is streamed to the device in chunks through two pinned staging
    buffers, so the host copy, the PCIe transfer and the decode of consecutive chunks overlap;
    pinned arrays and CPU tensors go over in one asynchronous copy.
    """
    rt.require_device()
    pageable = None
    if isinstance(heatmaps, np.ndarray) and heatmaps.ndim == 4 and heatmaps.dtype == np.float32 \
            and heatmaps.flags['C_CONTIGUOUS'] and heatmaps.flags['WRITEABLE'] \
            and heatmaps.nbytes >= PAGEABLE_MIN_BYTES and heatmaps.shape[0] % nviews == 0 \
            and heatmaps.shape[0] > 0 and not torch.from_numpy(heatmaps).is_pinned():
        pageable = heatmaps
        N, J, H, W = [int(v) for v in heatmaps.shape]
        views = [None]
    else:
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
    if pageable is not None:
        _lift_from_pageable(pageable, B, nviews, J, H, W, affine.view(N, 6), post_process, table,
                            no_distortion, conf_thre, (xy, maxvals, idx, poses3d, err, proj, resid), fmat, slots)
    else:
        _lift_launch(views, B, nviews, J, H, W, affine, post_process, table.pack, table.index, no_distortion,
                     conf_thre, xy, maxvals, idx, poses3d, err, proj, fmat, slots, resid, after_decode)
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
