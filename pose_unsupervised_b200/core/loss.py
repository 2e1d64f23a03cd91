"""Epipolar-consistency scoring behind the name of lib/core/loss.py:89-133.

``FundamentalLoss`` keeps the reference's call signature and returns the same scalar
(forward value; the autograd variant used while training is listed as a "next" row
in SURVEY.md section 8f).  ``epipolar_residuals`` is the per-(frame, pair, joint) form of
run/test/test_fund_mtx.py:56-69.  The reference reads {(subject, a, b): F} from
data/testdata/fundamental_matrix.pkl (lib/core/loss.py:92-94); here the dict is
passed in (or loaded from ``cfg.DATASET.ROOT`` when present).
"""
import itertools
import os
import pickle

import numpy as np
import torch

from .. import _lib, runtime as rt


class FundamentalTable(object):
    """{(subject, a, b): F[3,3]} packed as CUDA float64 [S, V, V, 9] + subject -> slot map."""

    def __init__(self, fdict, nviews=4):
        subjects = sorted({k[0] for k in fdict})
        self.slot = {s: i for i, s in enumerate(subjects)}
        self.nviews = nviews
        host = np.zeros((len(subjects), nviews, nviews, 9))
        for (s, a, b), F in fdict.items():
            if a < nviews and b < nviews:
                host[self.slot[s], a, b] = np.asarray(F, dtype=np.float64).reshape(9)
        # the reference indexes the dict per (subject, a, b) and raises KeyError on a missing pair
        # (lib/core/loss.py:123); a silent zero matrix would make residual and gradient too small
        missing = [(s, a, b) for s in subjects for a in range(nviews) for b in range(nviews)
                   if a != b and (s, a, b) not in fdict]
        if missing:
            raise KeyError('fundamental matrices missing for (subject, view a, view b): %s%s'
                           % (missing[:4], ' ...' if len(missing) > 4 else ''))
        self.fmat = rt.to_device(host)

    @classmethod
    def from_cameras(cls, cams_by_subject):
        """{subject: [V camera dicts]} -> exact F for every ordered view pair, computed on the GPU
        (replaces run/test/generate_fundamental_matirx.py, which fits F to data with LMedS)."""
        from ..multiviews.cameras import pack_camera
        rt.require_device()
        subjects = sorted(cams_by_subject)
        nviews = len(cams_by_subject[subjects[0]])
        pack = np.stack([pack_camera(c) for s in subjects for c in cams_by_subject[s]])
        ia, ib = [], []
        for si in range(len(subjects)):
            for a in range(nviews):
                for b in range(nviews):
                    ia.append(si * nviews + a)
                    ib.append(si * nviews + b)
        d_pack = rt.to_device(pack)
        d_a, d_b = rt.to_device(np.array(ia, dtype=np.int32)), rt.to_device(np.array(ib, dtype=np.int32))
        F = rt.empty((len(ia), 9), torch.float64)
        _lib.call('pb200_fundamental', rt.ptr(d_pack), rt.ptr(d_a), rt.ptr(d_b), len(ia), rt.ptr(F),
                  rt.stream_ptr())
        self = cls.__new__(cls)
        self.slot = {s: i for i, s in enumerate(subjects)}
        self.nviews = nviews
        self.fmat = F.view(len(subjects), nviews, nviews, 9).contiguous()
        return self

    def as_dict(self):
        """{(subject, a, b): F[3,3]} like the reference's fundamental_matrix.pkl (a != b)."""
        host = self.fmat.cpu().numpy()
        return {(s, a, b): host[i, a, b].reshape(3, 3) for s, i in self.slot.items()
                for a in range(self.nviews) for b in range(self.nviews) if a != b}

    def slots(self, subjects):
        """Subject ids [B] -> CUDA int32 table slots.  A CUDA int32 tensor is taken to hold slots
        already (compute them once with this method when the same frames are scored repeatedly)."""
        if isinstance(subjects, torch.Tensor) and subjects.is_cuda and subjects.dtype == torch.int32:
            return subjects
        subs = np.asarray(subjects.cpu() if isinstance(subjects, torch.Tensor) else subjects).reshape(-1)
        keys = np.array(sorted(self.slot))
        pos = np.searchsorted(keys, subs)
        if np.any(pos >= len(keys)) or np.any(keys[np.minimum(pos, len(keys) - 1)] != subs):
            raise KeyError('subject without fundamental matrices')
        lut = np.array([self.slot[k] for k in keys.tolist()], dtype=np.int32)
        return rt.to_device(lut[pos])


def epipolar_residuals(pred2d, subjects, fundamental, nviews=4, weight=None, return_sum=False):
    """|x_b^T F_(subj,a,b) x_a| for the V(V-1) ordered pairs of every frame.

    pred2d [B*V, J, 2] view-minor (numpy or CUDA, float32/float64); subjects [B];
    fundamental: FundamentalTable or the reference's dict.  Returns [B, V(V-1), J] float64
    (pairs in itertools.permutations order), optionally with the grand total.
    """
    rt.require_device()
    table = fundamental if isinstance(fundamental, FundamentalTable) else FundamentalTable(fundamental, nviews)
    xy = rt.to_device_float(pred2d)
    N, J = int(xy.shape[0]), int(xy.shape[1])
    if N % nviews != 0:
        raise ValueError('%d rows are not a multiple of nviews=%d' % (N, nviews))
    if table.nviews != nviews:
        raise ValueError('the fundamental table holds %d views, nviews=%d' % (table.nviews, nviews))
    B = N // nviews
    subj = table.slots(subjects)
    assert int(subj.shape[0]) == B, 'one subject per frame'
    w = rt.to_device_float(weight).reshape(N, J) if weight is not None else None
    resid = rt.empty((B, nviews * (nviews - 1), J), torch.float64)
    total = rt.zeros((1,), torch.float64) if return_sum else None
    _lib.call('pb200_epipolar', rt.ptr(table.fmat), rt.ptr(subj), rt.ptr(xy), rt.float_dtype_tag(xy),
              rt.ptr(w), rt.float_dtype_tag(w) if w is not None else _lib.F64, B, nviews, J,
              rt.ptr(resid), rt.ptr(total), rt.stream_ptr())
    if not rt.is_device_tensor(pred2d):
        resid = rt.to_host(resid)
    return (resid, total) if return_sum else resid


class _EpipolarLossFn(torch.autograd.Function):
    """sum over frames, ordered pairs and joints of |x_b^T F x_a| (* w_b * w_a), times `scale`."""

    @staticmethod
    def forward(ctx, xy, weight, subj_slots, fmat, nviews, scale):
        x = xy.detach().contiguous()
        N, J = int(x.shape[0]), int(x.shape[1])
        B = N // nviews
        w = None if weight is None else weight.detach().reshape(N, J).contiguous()
        resid = rt.empty((B, nviews * (nviews - 1), J), torch.float64)
        total = rt.zeros((1,), torch.float64)
        _lib.call('pb200_epipolar', rt.ptr(fmat), rt.ptr(subj_slots), rt.ptr(x), rt.float_dtype_tag(x),
                  rt.ptr(w), rt.float_dtype_tag(w) if w is not None else _lib.F64, B, nviews, J,
                  rt.ptr(resid), rt.ptr(total), rt.stream_ptr())
        ctx.save_for_backward(x, w if w is not None else x.new_empty(0), subj_slots, fmat)
        ctx.has_w, ctx.nviews, ctx.scale = w is not None, nviews, float(scale)
        return (total[0] * scale).to(xy.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        x, w, subj_slots, fmat = ctx.saved_tensors
        N, J = int(x.shape[0]), int(x.shape[1])
        B = N // ctx.nviews
        g = rt.zeros((N, J, 2), torch.float64)
        _lib.call('pb200_epipolar_grad', rt.ptr(fmat), rt.ptr(subj_slots), rt.ptr(x), rt.float_dtype_tag(x),
                  rt.ptr(w) if ctx.has_w else None, rt.float_dtype_tag(w) if ctx.has_w else _lib.F64,
                  B, ctx.nviews, J, ctx.scale, rt.ptr(g), rt.stream_ptr())
        return (g * grad_out.to(torch.float64)).to(x.dtype), None, None, None, None, None


class FundamentalLoss(object):
    """lib/core/loss.py:89-133, forward and backward (the gradient flows to ``joints_2d_list``;
    ``target_weight`` is treated as a constant, as in the reference's use)."""

    def __init__(self, cfg, fundamental_matrix_dict=None):
        self.use_target_weight = cfg.LOSS.USE_TARGET_WEIGHT_FUND
        if fundamental_matrix_dict is None:
            path = os.path.join(cfg.DATASET.ROOT, 'testdata', 'fundamental_matrix.pkl')
            with open(path, 'rb') as f:
                fundamental_matrix_dict = pickle.load(f)
        self.fundamental_matrix_dict = fundamental_matrix_dict
        self._tables = {}

    def __call__(self, joints_2d_list, target_weight, meta):
        """joints_2d_list: V tensors [K,J,2] (image px); target_weight: V tensors [K,J,1];
        meta: V dicts with 'subject' [K].  Returns a 0-d CUDA float64 tensor."""
        assert isinstance(joints_2d_list[0], torch.Tensor)
        nviews = len(joints_2d_list)
        K, J = joints_2d_list[0].shape[:2]
        subject = meta[0]['subject']
        subject = subject.numpy() if isinstance(subject, torch.Tensor) else np.asarray(subject)
        assert K == len(subject)
        table = self._tables.get(nviews)
        if table is None:
            table = self._tables[nviews] = FundamentalTable(self.fundamental_matrix_dict, nviews)
        # view-minor rows: row = sample * V + view (torch.stack keeps the autograd graph)
        xy = torch.stack(list(joints_2d_list), dim=1).reshape(K * nviews, J, 2)
        if not xy.is_cuda:
            xy = xy.to(rt.device())
        w = None
        if self.use_target_weight:
            w = torch.stack([t.detach() for t in target_weight], dim=1).reshape(K * nviews, J).to(xy.device)
            if w.dtype not in (torch.float32, torch.float64):
                w = w.to(torch.float32)
        npairs = len(list(itertools.permutations(range(nviews), 2)))
        return _EpipolarLossFn.apply(xy, w, table.slots(subject), table.fmat, nviews, 1.0 / (K * npairs * J))
