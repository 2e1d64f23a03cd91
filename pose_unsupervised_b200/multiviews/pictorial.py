"""Recursive pictorial structure model behind the names of lib/multiviews/pictorial.py.

``rpsm`` keeps the reference's signature (lib/multiviews/pictorial.py:214-250) for one
frame; ``rpsm_batch`` runs a batch of frames, one thread block per frame
(csrc/rpsm.cu).  The level-0 ``pairwise_constraint`` dict of scipy sparse matrices
({(parent, child): [nbins, nbins]}, run/test/test_rpsm.py:131-135) is converted once
into a device bit matrix (``PairwiseTable``) and cached; ``PairwiseTable.from_limb_lengths``
builds it on the GPU from average limb lengths instead of the reference's offline
O(n^6) Python generator (run/test/generate_pairwise_constraints.py:60-95).
"""
import numpy as np
import torch

from .. import _lib, runtime as rt
from ..utils.transforms import crop_affine
from .body import HumanBody
from .cameras import CameraTable


def pack_pairwise_bits(m):
    """Dense 0/1 matrix [n, n] -> uint32 [n, ceil(n/32)]: bit (j % 32) of word [i, j // 32] is P[i, j]
    (the layout pb200_rpsm reads; host-side packing of the reference's scipy matrices)."""
    m = np.asarray(m)
    nb = m.shape[0]
    if m.shape != (nb, nb):
        raise ValueError('pairwise matrix must be square')
    if not np.all((m == 0) | (m == 1)):
        raise ValueError('pairwise matrix must hold only 0/1 (generate_pairwise_constraints.py:93)')
    words = (nb + 31) // 32
    padded = np.zeros((nb, words * 32), dtype=np.uint8)
    padded[:, :nb] = m != 0
    return np.packbits(padded, axis=1, bitorder='little').view('<u4').astype(np.uint32)


class PairwiseTable(object):
    """Level-0 pairwise constraints as bits: [E, nbins, ceil(nbins/32)] uint32 on the device."""

    _cache = {}

    def __init__(self, bits, nbins):
        self.bits = bits
        self.nbins = nbins
        self._offset_only = None
        self._max_reach = -1

    def _check(self):
        n = int(round(self.nbins ** (1.0 / 3)))
        if n ** 3 != self.nbins:
            self._offset_only, self._max_reach = False, -1
        else:
            flag = rt.zeros((2,), torch.int32)
            _lib.call('pb200_pairwise_lut_check', rt.ptr(self.bits), int(self.bits.shape[0]), n,
                      rt.ptr(flag), rt.stream_ptr())
            bad, reach = [int(v) for v in flag.cpu()]
            self._offset_only, self._max_reach = bad == 0, reach

    @property
    def offset_only(self):
        """True if P[i,j] depends on the index offset (|dy|,|dx|,|dz|) only (checked once, on the GPU)."""
        if self._offset_only is None:
            self._check()
        return self._offset_only

    @property
    def max_reach(self):
        """Largest |bin offset| of an allowed pair in row 0 of any edge (decides, with ``offset_only``,
        whether level 0 can run on chip; checked once, on the GPU)."""
        if self._offset_only is None:
            self._check()
        return self._max_reach

    @classmethod
    def from_dict(cls, pairwise, body):
        """{(parent, child): sparse or dense [nbins, nbins]} -> table (cached per dict object)."""
        if isinstance(pairwise, PairwiseTable):
            return pairwise
        key = (id(pairwise), torch.cuda.current_device())
        hit = cls._cache.get(key)
        if hit is not None and hit[0] is pairwise:
            return hit[1]
        rows = []
        for e in body.edges():
            m = pairwise[e]
            m = m.toarray() if hasattr(m, 'toarray') else np.asarray(m)
            rows.append(pack_pairwise_bits(m))
        host = np.stack(rows)
        table = cls(rt.to_device(host.view(np.int32)), host.shape[1])
        if len(cls._cache) > 4:
            cls._cache.clear()
        cls._cache[key] = (pairwise, table)
        return table

    @classmethod
    def from_limb_lengths(cls, avg_limb_length, body, box_size=2000, nbins=16):
        """P[i,j] = | |g_i - g_j| - L | < 0.4 L on the zero-centred grid, built on the GPU."""
        rt.require_device()
        edges = body.edges()
        limb = rt.to_device(np.array([avg_limb_length[e] for e in edges], dtype=np.float64))
        nb = nbins ** 3
        bits = rt.empty((len(edges), nb, (nb + 31) // 32), torch.int32)
        _lib.call('pb200_pairwise_level0', rt.ptr(limb), len(edges), int(nbins), float(box_size),
                  rt.ptr(bits), rt.stream_ptr())
        return cls(bits, nb)

    def to_dense(self, edge_index):
        """Host bool [nbins, nbins] of one edge (tests, inspection)."""
        w = self.bits[edge_index].cpu().numpy().view(np.uint32)
        bits = np.unpackbits(w.view(np.uint8), axis=1, bitorder='little')
        return bits[:, :self.nbins].astype(bool)


def rpsm_batch(cams, heatmaps, boxes_center, boxes_scale, grid_centers, limb_lengths, pairwise,
               config, body=None, return_trace=False, use_lut=True, onchip=True):
    """RPSM for B frames.

    cams         : CameraTable or list of B*V camera dicts (view-minor rows)
    heatmaps     : [B, V, J, H, W] float32 (numpy or CUDA)
    boxes_center : [B*V, 2], boxes_scale [B*V, 2] crop boxes the heatmaps were computed on
    grid_centers : [B, 3] float64 root locations
    limb_lengths : [B, E] float64 in ``body.edges()`` order
    pairwise     : PairwiseTable or the reference's dict
    use_lut      : allow the shared-memory offset table when the pairwise matrix permits it
    onchip       : allow the on-chip level 0 (csrc/rpsm.cu::rpsm_onchip_kernel) when the table and the
                   grid permit it; False forces the generic kernel (same results, used by the tests)
    Returns poses [B, J, 3] float64 (CUDA tensor if heatmaps is one), optionally the chosen
    bins per level [B, depth+1, J] int32.
    """
    rt.require_device()
    body = HumanBody() if body is None else body
    hm = rt.to_device(heatmaps)
    if hm.dtype != torch.float32 or hm.dim() != 5:
        raise TypeError('heatmaps must be float32 [B, V, J, H, W]')
    B, V, J, H, W = [int(v) for v in hm.shape]
    if J != len(body.skeleton):
        raise ValueError('heatmaps have %d joints, the body %d' % (J, len(body.skeleton)))
    if H != W:
        # the reference builds its interpolator on (arange(h), arange(w)) with hmap.T,
        # which only works for square maps (lib/multiviews/pictorial.py:176-186)
        raise ValueError('RPSM needs square heatmaps')
    table = CameraTable.from_cameras(cams)
    ps = config.PICT_STRUCT
    img = config.NETWORK.IMAGE_SIZE
    aff = crop_affine(boxes_center, boxes_scale, (int(img[0]), int(img[1])), inv=0)
    if aff.shape[0] != B * V or len(table) < B * V:
        raise ValueError('need one camera and one box per (frame, view) row')
    edges, order, root_idx = body.tree_arrays()
    E = len(edges)
    root = rt.to_device(np.asarray(grid_centers, dtype=np.float64).reshape(B, 3)
                        if not isinstance(grid_centers, torch.Tensor) else grid_centers, torch.float64)
    limb = rt.to_device(np.asarray(limb_lengths, dtype=np.float64).reshape(B, E)
                        if not isinstance(limb_lengths, torch.Tensor) else limb_lengths, torch.float64)
    pw = PairwiseTable.from_dict(pairwise, body)
    n0 = int(ps.FIRST_NBINS)
    if pw.nbins != n0 ** 3 or pw.bits.shape[0] != E:
        raise ValueError('pairwise table is %s, expected [%d, %d, .]' % (tuple(pw.bits.shape), E, n0 ** 3))
    lib = _lib.load()
    nbytes = lib.pb200_rpsm_workspace_bytes(B, J, n0, lib.pb200_sm_count())
    ws = rt.workspace('rpsm', nbytes)
    d_edges, d_order = rt.to_device(edges), rt.to_device(order)
    depth = int(ps.RECUR_DEPTH)
    pose = rt.empty((B, J, 3), torch.float64)
    trace = rt.empty((B, depth + 1, J), torch.int32) if return_trace else None
    _lib.call('pb200_rpsm', rt.ptr(hm), B, V, J, H, W, rt.ptr(table.pack), rt.ptr(table.index),
              rt.ptr(aff), int(img[0]), int(img[1]), rt.ptr(root), rt.ptr(limb),
              rt.ptr(d_edges), rt.ptr(d_order), root_idx, rt.ptr(pw.bits), int(pw.offset_only and use_lut),
              int(pw.max_reach) if (onchip and use_lut and pw.offset_only) else -1, n0, int(ps.RECUR_NBINS), depth, float(ps.GRID_SIZE), float(ps.LIMB_LENGTH_TOLERANCE),
              rt.ptr(ws), int(ws.numel()), rt.ptr(pose), rt.ptr(trace), rt.stream_ptr())
    if not rt.is_device_tensor(heatmaps):
        pose = rt.to_host(pose)
        trace = rt.to_host(trace) if return_trace else None
    return (pose, trace) if return_trace else pose


def rpsm(cams, heatmaps, boxes, grid_center, limb_length, pairwise_constraint, config, body=None):
    """lib/multiviews/pictorial.py:214-250 for one frame -> pose3d [J,3] float64.

    cams: V camera dicts; heatmaps [V,J,H,W]; boxes: V dicts {center, scale};
    limb_length {(parent, child): mm}; pairwise_constraint: the reference's dict (or a
    PairwiseTable).
    """
    body = HumanBody() if body is None else body
    hm = np.asarray(heatmaps)
    # keep float32 boxes float32: `scale * 200.0` is rounded in the dtype of `scale`
    # (lib/utils/transforms.py:84)
    def stack(key):
        vals = [np.asarray(b[key]).reshape(-1)[:2] for b in boxes]
        f32 = all(v.dtype == np.float32 for v in vals)
        return np.array(vals, dtype=np.float32 if f32 else np.float64)
    centers, scales = stack('center'), stack('scale')
    limb = np.array([limb_length[e] for e in body.edges()], dtype=np.float64)[None]
    return rpsm_batch(list(cams), hm[None], centers, scales,
                      np.asarray(grid_center, dtype=np.float64).reshape(1, 3), limb,
                      pairwise_constraint, config, body)[0]


def break_limb_length(poses3d, limb_length, body=None, thres=0.4):
    """run/pose3d/estimate.py:84-96 for a batch: flag [B] uint8, 1 where some limb of the pose
    deviates from its expected length by more than ``thres`` x expected.

    limb_length: [E] (one template) or [B,E] in ``body.edges()`` order, or the reference's dict.
    """
    rt.require_device()
    body = HumanBody() if body is None else body
    edges, _, _ = body.tree_arrays()
    p = rt.to_device(poses3d, torch.float64)
    B, J = int(p.shape[0]), int(p.shape[1])
    if isinstance(limb_length, dict):
        limb_length = np.array([limb_length[e] for e in body.edges()], dtype=np.float64)
    L = rt.to_device(limb_length, torch.float64)
    per_frame = int(L.dim() == 2)
    flag = rt.empty((B,), torch.uint8)
    _lib.call('pb200_limb_break', rt.ptr(p), rt.ptr(rt.to_device(edges)), rt.ptr(L), per_frame, B, J,
              len(edges), float(thres), rt.ptr(flag), rt.stream_ptr())
    return flag if rt.is_device_tensor(poses3d) else rt.to_host(flag)


def lift_combination(cams, heatmaps, boxes_center, boxes_scale, poses2d, limb_lengths, pairwise, config,
                     body=None, joints_vis=None, thres=0.4):
    """The "combination" estimator sketched in run/pose3d/estimate.py:229-244: triangulate every
    frame; where a limb of the triangulated pose breaks its expected length by more than ``thres``
    the frame is re-estimated by RPSM with the triangulated root joint as grid centre.

    heatmaps [B,V,J,H,W]; poses2d [B*V,J,2] (e.g. decoded from the same heatmaps); limb_lengths
    [B,E] or [E].  Returns (poses3d [B,J,3] float64 numpy, used_rpsm [B] bool numpy).
    """
    from .triangulate import triangulate_poses
    body = HumanBody() if body is None else body
    table = CameraTable.from_cameras(cams)
    hm = rt.to_device(heatmaps)
    B, V = int(hm.shape[0]), int(hm.shape[1])
    poses = triangulate_poses(table, rt.to_device_float(poses2d),
                              None if joints_vis is None else rt.to_device(joints_vis), nviews=V)
    L = rt.to_device(np.asarray(limb_lengths, dtype=np.float64) if not isinstance(limb_lengths, torch.Tensor)
                     else limb_lengths, torch.float64)
    flag = break_limb_length(poses, L, body, thres).bool()
    sel = torch.nonzero(flag).reshape(-1)
    if sel.numel() > 0:
        rows = (sel[:, None] * V + torch.arange(V, device=sel.device)[None]).reshape(-1)
        sub_table = CameraTable(table.pack, table.index[rows].contiguous())
        c = rt.to_device_float(boxes_center).reshape(B * V, 2)[rows]
        s = rt.to_device_float(boxes_scale).reshape(B * V, 2)[rows]
        limb_sel = L[sel] if L.dim() == 2 else L[None].expand(sel.numel(), -1).contiguous()
        fixed = rpsm_batch(sub_table, hm[sel].contiguous(), c, s, poses[sel, body.root_idx].contiguous(),
                           limb_sel, pairwise, config, body)
        poses = poses.clone()
        poses[sel] = fixed
    return rt.to_host(poses), rt.to_host(flag)
