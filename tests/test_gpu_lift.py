"""GPU parity of the fused pass (decode -> triangulate -> reprojection error, one kernel)
against decode + reproject_poses of the oracle, and against the unfused CUDA entry points."""
import numpy as np
import pytest
import torch

from oracle import inference as oinf
from oracle import triangulate as otri
from pose_unsupervised_b200.utils import synth
from tests.util import ulp_diff_f32

pytestmark = pytest.mark.gpu


def _inputs(B, V, J, hw, seed=0, peaked=True):
    rng = np.random.default_rng(seed)
    rigs = synth.camera_table(3, V, seed=seed)
    subj = rng.integers(0, 3, B)
    cams = [rigs[s][v] for s in subj for v in range(V)]
    hm = rng.random((B * V, J, hw, hw), dtype=np.float32) * 0.5
    center = rng.uniform(400, 600, (B * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    if peaked:
        # consistent peaks: project a pose, map it into each crop, stamp a maximum there
        poses = synth.random_poses(B, seed=seed + 1, njoints=J)
        for i in range(B):
            for v in range(V):
                r = i * V + v
                xy = synth.project_plumb_bob_numpy(poses[i], cams[r])
                t = synth.crop_affine_numpy(center[r], scale[r, 0], hw, hw)
                px = np.clip(np.round(xy @ t[:, :2].T + t[:, 2]), 0, hw - 1).astype(int)
                hm[r, np.arange(J), px[:, 1], px[:, 0]] = 1.0 + rng.random(J).astype(np.float32)
    return hm, center, scale, cams


@pytest.mark.parametrize('V,J,hw', [(4, 17, 64), (2, 16, 64), (8, 17, 64), (4, 5, 96), (4, 17, 48)])
def test_fused_vs_oracle(V, J, hw):
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    B = 24
    hm, center, scale, cams = _inputs(B, V, J, hw, seed=V + J)
    res = lift_heatmaps(hm, center, scale, cams, nviews=V, post_process=True, return_idx=True,
                        return_proj=True).numpy()
    rp, rm = oinf.get_final_preds(True, hm, center, scale)
    assert np.array_equal(res.idx, oinf.flat_argmax(hm))
    assert np.array_equal(res.maxvals, rm[:, :, 0])
    assert ulp_diff_f32(res.xy, rp).max() <= 1
    vis = np.ones(rp.shape[:2])
    # the oracle lifts the coordinates the kernel decoded (float32), like the h5 hand-off
    proj, rvis, pts = otri.reproject_poses(res.xy, cams, vis, nviews=V, return_points=True)
    assert np.abs(res.poses3d - pts).max() < 1e-2
    assert np.abs(res.proj2d - proj.astype(np.float64)).max() < 1e-2      # oracle proj is float32 here
    err = np.linalg.norm(res.proj2d - res.xy.astype(np.float64), axis=2)
    assert np.abs(res.reproj_err - err).max() < 1e-3 * max(1.0, err.max())


def test_fused_confidence_threshold_and_unfused_agree():
    from pose_unsupervised_b200.core.inference import decode_heatmaps
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps, reproject_poses
    B, V, J = 40, 4, 17
    hm, center, scale, cams = _inputs(B, V, J, 64, seed=5)
    hm[::5] *= 0.3                                                    # some rows below the threshold
    hm[:12] *= 0.3                                                    # and three whole frames
    thre = 0.6
    res = lift_heatmaps(hm, center, scale, cams, conf_thre=thre, return_proj=True)
    xy, mv = decode_heatmaps(hm, center, scale, post_process=True)
    assert torch.equal(res.xy, xy) and torch.equal(res.maxvals, mv)
    vis = (mv > thre)
    proj, pvis, pts = reproject_poses(xy, cams, vis, return_points=True)
    assert torch.equal(res.poses3d, pts)
    assert torch.equal(res.proj2d.float(), proj)
    ovis = mv.cpu().numpy() > thre
    ref = otri.triangulate_poses(cams, xy.cpu().numpy(), ovis)
    assert np.abs(res.poses3d.cpu().numpy() - ref).max() < 1e-2
    assert (ref == 0).all(axis=2).any()                               # some joints had < 2 views


def test_lift_shapes_nan_and_thresholds():
    """Lift = decode + per-joint lift: identical bits to the stand-alone entries, incl. NaN maps and 80x80."""
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    for hw in (64, 80, 32):
        hm, center, scale, cams = _inputs(16, 4, 17, hw, seed=hw)
        hm[3, 2, 5, 7] = np.nan
        hm[9, 0] = -1.0
        res = lift_heatmaps(hm, center, scale, cams, return_idx=True).numpy()
        thr = lift_heatmaps(hm, center, scale, cams, conf_thre=1.5, return_proj=True).numpy()
        few = (thr.maxvals.reshape(16, 4, 17) > 1.5).sum(axis=1) < 2          # joints with < 2 confident views
        assert few.any() and not few.all()
        assert np.all(thr.poses3d[few] == 0) and np.all(thr.reproj_err.reshape(16, 4, 17)[:, 0][few] == 0)
        vis = np.nan_to_num(thr.maxvals, nan=-1.0) > 1.5
        ref_thr = otri.triangulate_poses(cams, np.nan_to_num(thr.xy), vis)
        assert np.abs(thr.poses3d - ref_thr)[~few].max() < 1e-2
        assert np.array_equal(res.idx, oinf.flat_argmax(hm)), hw
        rp, rm = oinf.get_final_preds(True, hm, center, scale)
        assert np.array_equal(res.maxvals, rm[:, :, 0], equal_nan=True)
        assert ulp_diff_f32(res.xy, rp).max() <= 1
        ok = ~np.isnan(res.xy).any(axis=2)
        pts = otri.triangulate_poses(cams, np.nan_to_num(res.xy), ok)
        good = ok.reshape(16, 4, 17).all(axis=1)
        assert np.abs(res.poses3d - pts)[good].max() < 1e-2


def test_lift_repeated_launches_and_view_lists():
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    hm, center, scale, cams = _inputs(64, 4, 17, 64, seed=9)
    d_hm = torch.from_numpy(hm).cuda()
    first = lift_heatmaps(d_hm, center, scale, cams)
    for _ in range(5):
        again = lift_heatmaps(d_hm, center, scale, cams)
        assert torch.equal(first.poses3d, again.poses3d) and torch.equal(first.reproj_err, again.reproj_err)
    views = [d_hm.view(64, 4, 17, 64, 64)[:, v].contiguous() for v in range(4)]
    listed = lift_heatmaps(views, center, scale, cams)
    assert torch.equal(first.poses3d, listed.poses3d) and torch.equal(first.xy, listed.xy)


def test_fused_full_size_properties():
    """BASELINE.json config 2: 4096 frames x 4 views x 17 joints x 64x64 on one GPU."""
    from pose_unsupervised_b200.core.inference import decode_heatmaps
    from pose_unsupervised_b200.multiviews.cameras import CameraTable, pack_camera
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps, triangulate_poses
    B, V, J = 4096, 4, 17
    g = torch.Generator(device='cuda').manual_seed(1)
    hm = torch.rand((B * V, J, 64, 64), generator=g, device='cuda')
    rng = np.random.default_rng(2)
    rigs = synth.camera_table(7, 4, seed=3)
    pack = np.array([pack_camera(c) for rig in rigs for c in rig])
    subj = rng.integers(0, 7, B)
    table = CameraTable.from_arrays(pack, (subj[:, None] * 4 + np.arange(4)[None]).reshape(-1))
    center = rng.uniform(400, 600, (B * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    res = lift_heatmaps(hm, center, scale, table, return_idx=True)
    xy, mv, idx = decode_heatmaps(hm, center, scale, post_process=True, return_idx=True)
    assert torch.equal(res.idx, idx) and torch.equal(res.xy, xy) and torch.equal(res.maxvals, mv)
    assert torch.equal(res.poses3d, triangulate_poses(table, xy))
    sl = slice(777, 793)
    cams = [rigs[subj[i]][v] for i in range(sl.start, sl.stop) for v in range(4)]
    ref = otri.triangulate_poses(cams, xy[sl.start * 4:sl.stop * 4].cpu().numpy())
    got = res.poses3d[sl].cpu().numpy()
    # random heatmaps give inconsistent views: the DLT solution is far away and ill-conditioned,
    # so compare relative to the magnitude of the point
    assert (np.abs(got - ref) / np.maximum(1.0, np.abs(ref))).max() < 1e-6


def test_lift_with_epipolar_residuals_in_the_same_pass():
    """north_star: reprojection error and epipolar residuals come out of the same pass.  They must equal
    the stand-alone epipolar kernel on the decoded coordinates, and the oracle."""
    from oracle import epipolar as oepi
    from pose_unsupervised_b200 import _lib
    from pose_unsupervised_b200.core.loss import FundamentalTable, epipolar_residuals
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    B, V, J = 12, 4, 17
    rng = np.random.default_rng(3)
    rigs = synth.camera_table(3, V, seed=3)
    subj = rng.integers(0, 3, B)
    cams = [rigs[s][v] for s in subj for v in range(V)]
    hm = rng.random((B * V, J, 64, 64), dtype=np.float32)
    center = rng.uniform(400, 600, (B * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    table = FundamentalTable.from_cameras({s: rigs[s] for s in range(3)})
    res = lift_heatmaps(hm, center, scale, cams, fundamental=table, subjects=subj)
    alone = epipolar_residuals(res.xy, subj, table)
    assert res.epipolar.shape == (B, V * (V - 1), J) and torch.equal(res.epipolar, alone)
    F = oepi.fundamental_table({s: rigs[s] for s in range(3)})
    ref = oepi.epipolar_residuals(res.xy.cpu().numpy().astype(np.float64), subj, F)
    assert np.abs(res.epipolar.cpu().numpy() - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())


def test_pageable_host_pipeline_matches_the_plain_copy():
    """numpy heatmaps in pageable memory go through the chunked pinned-staging pipeline
    (odd chunk count, a short last chunk, two back-to-back calls reusing the staging buffers):
    every output equals the device-resident call bit for bit."""
    from pose_unsupervised_b200.core.loss import FundamentalTable
    from pose_unsupervised_b200.multiviews import triangulate as tri
    B, V, J = 37, 4, 17
    rng = np.random.default_rng(21)
    rigs = synth.camera_table(3, V, seed=3)
    subj = rng.integers(0, 3, B)
    cams = [rigs[s][v] for s in subj for v in range(V)]
    hm = rng.random((B * V, J, 64, 64), dtype=np.float32)
    center = rng.uniform(400, 600, (B * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    table = FundamentalTable.from_cameras({s: rigs[s] for s in range(3)})
    ref = tri.lift_heatmaps(torch.from_numpy(hm).cuda(), center, scale, cams, conf_thre=0.2, return_idx=True,
                            return_proj=True, fundamental=table, subjects=subj)
    old = tri.STAGE_BYTES, tri.PAGEABLE_MIN_BYTES
    tri.STAGE_BYTES, tri.PAGEABLE_MIN_BYTES = 5 * V * J * 64 * 64 * 4, 0          # 5 frames per chunk -> 8 chunks
    try:
        for _ in range(2):
            got = tri.lift_heatmaps(hm, center, scale, cams, conf_thre=0.2, return_idx=True, return_proj=True,
                                    fundamental=table, subjects=subj)
            for name in ('xy', 'maxvals', 'idx', 'poses3d', 'reproj_err', 'proj2d', 'epipolar'):
                assert torch.equal(getattr(got, name), getattr(ref, name)), name
    finally:
        tri.STAGE_BYTES, tri.PAGEABLE_MIN_BYTES = old


def test_full_size_launches_are_repeatable_and_schedule_independent():
    """compute-sanitizer is closed on the pool, so race freedom of the hand-offs is checked the blunt way: the
    BASELINE batch (4096 frames x 4 views x 17 joints x 64^2) lifted five times, with both decode map
    schedules (static striding / strided share + claimed tail), must give identical bits every time -- a lost
    update or a map decoded twice / not at all would show up as a differing or stale entry."""
    from pose_unsupervised_b200 import runtime as rt
    from pose_unsupervised_b200.multiviews.cameras import CameraTable, pack_camera
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    B, V, J = 4096, 4, 17
    g = torch.Generator(device='cuda').manual_seed(7)
    hm = torch.rand((B * V, J, 64, 64), generator=g, device='cuda', dtype=torch.float32)
    rng = np.random.default_rng(7)
    rigs = synth.camera_table(7, V, seed=0)
    pack = np.array([pack_camera(c) for rig in rigs for c in rig])
    subj = rng.integers(0, 7, B)
    table = CameraTable.from_arrays(pack, (subj[:, None] * V + np.arange(V)[None]).reshape(-1))
    center = rng.uniform(400, 600, (B * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    try:
        ref = None
        for trial in range(5):
            rt.set_decode_schedule(trial % 2 == 0)
            xy_poison = None
            res = lift_heatmaps(hm, center, scale, table, return_idx=True)
            got = (res.idx.clone(), res.maxvals.clone(), res.xy.clone(), res.poses3d.clone(), res.reproj_err.clone())
            if ref is None:
                ref = got
                assert torch.equal(ref[0].long(), hm.view(B * V, J, -1).argmax(dim=2))
            else:
                for a, b in zip(ref, got):
                    assert torch.equal(a, b), trial
    finally:
        rt.set_decode_schedule(True)
