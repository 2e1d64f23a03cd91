"""CPU tests: the oracle against the golden vectors recorded from the real reference
(tests/golden/make_golden.py) and against self-made known-answer tests where the
reference has none (the pymvg boundary).  No GPU, no /root/reference at run time."""
import itertools

import numpy as np
import pytest

from oracle import cameras as ocam
from oracle import epipolar as oepi
from oracle import inference as oinf
from oracle import pictorial as opict
from oracle import transforms as otr
from oracle import triangulate as otri
from oracle.body import HumanBody, h36m17
from oracle.pymvg_restated import RestatedCamera
from pose_unsupervised_b200.utils import synth
from tests.util import golden, rpsm_config, rpsm_golden_frame, unpack_cam


# ---- transforms -------------------------------------------------------------------
def test_affine_bit_exact_vs_reference():
    a = golden('affine.npz')
    for i in range(len(a['center'])):
        c, s = a['center'][i], a['scale'][i]
        if a['f32'][i]:
            c, s = c.astype(np.float32), s.astype(np.float32)
        fwd = otr.get_affine_transform(c, s, a['rot'][i], a['size'][i])
        inv = otr.get_affine_transform(c, s, a['rot'][i], a['size'][i], inv=1)
        assert np.array_equal(fwd, a['fwd'][i]) and np.array_equal(inv, a['inv'][i]), i


def test_affine_backends_agree():
    cv2 = pytest.importorskip('cv2')  # noqa: F841
    rng = np.random.default_rng(0)
    try:
        for _ in range(50):
            c, s = rng.uniform(100, 900, 2), np.repeat(rng.uniform(0.5, 4), 2)
            otr.set_backend('lu')
            a = otr.get_affine_transform(c, s, 0, [64, 64], inv=1)
            otr.set_backend('cv2')
            b = otr.get_affine_transform(c, s, 0, [64, 64], inv=1)
            assert np.array_equal(a, b)
    finally:
        otr.set_backend('lu')


# ---- decode -------------------------------------------------------------------------
def test_decode_matches_reference_golden():
    d = golden('decode.npz')
    for name in d['names']:
        hm, c, s = d[name + '_hm'], d[name + '_center'], d[name + '_scale']
        preds, maxvals = oinf.get_max_preds(hm)
        assert np.array_equal(preds, d[name + '_preds'])
        assert np.array_equal(maxvals, d[name + '_maxvals'], equal_nan=True)
        assert np.array_equal(oinf.flat_argmax(hm), d[name + '_idx'])
        for pp in (0, 1):
            with np.errstate(invalid='ignore'):
                vec, _ = oinf.get_final_preds(pp, hm, c, s)
                loops, _ = oinf.get_final_preds_loops(pp, hm, c, s)
            assert np.array_equal(vec, d[name + '_final%d' % pp], equal_nan=True), (name, pp)
            assert np.array_equal(loops, d[name + '_final%d' % pp], equal_nan=True), (name, pp)


def test_argmax_tie_rules():
    hm = np.zeros((1, 4, 4, 4), np.float32)
    hm[0, 0, 1, 2] = hm[0, 0, 3, 3] = 1.0           # first maximum wins
    hm[0, 1] = -1.0                                   # all <= 0 -> coordinates masked to 0
    hm[0, 2, 2, 1] = np.nan
    hm[0, 2, 0, 3] = 7.0                              # NaN beats any number
    hm[0, 3, 3, 0] = np.inf
    preds, maxvals = oinf.get_max_preds(hm)
    assert oinf.flat_argmax(hm).tolist() == [[6, 0, 9, 12]]
    assert preds[0, 0].tolist() == [2.0, 1.0] and preds[0, 1].tolist() == [0.0, 0.0]
    assert np.isnan(maxvals[0, 2, 0]) and preds[0, 2].tolist() == [0.0, 0.0]
    assert preds[0, 3].tolist() == [0.0, 3.0]


# ---- cameras ----------------------------------------------------------------------
def test_cameras_bit_exact_vs_reference():
    c = golden('cameras.npz')
    for i, v in enumerate(c['cams']):
        cam = unpack_cam(v)
        assert np.array_equal(ocam.project_pose(c['pts'], cam), c['proj'][i])
        assert np.array_equal(ocam.world_to_camera_frame(c['pts'], cam['R'], cam['T']), c['w2c'][i])
        assert np.array_equal(ocam.camera_to_world_frame(c['w2c'][i], cam['R'], cam['T']), c['c2w'][i])


# ---- pymvg boundary: self-made known-answer tests ----------------------------------
def _rig_and_obs(nviews, nframes, noise, distorted, seed=0):
    rigs = synth.camera_table(3, nviews, seed=seed)
    poses = synth.random_poses(nframes, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    obs, cams = synth.multiview_observations(poses, rigs, rng.integers(0, 3, nframes), noise_px=noise,
                                             seed=seed + 3, distorted=distorted)
    return poses, obs, cams


@pytest.mark.parametrize('nviews', [2, 4, 8])
def test_triangulate_noise_free_round_trip(nviews):
    poses, obs, cams = _rig_and_obs(nviews, 6, 0.0, distorted=False)
    X = otri.triangulate_poses(cams, obs, None, no_distortion=True, nviews=nviews)
    assert np.abs(X - poses).max() < 1e-6          # mm
    # with lens distortion the 5-iteration undistortion leaves a small residual
    poses, obs, cams = _rig_and_obs(nviews, 6, 0.0, distorted=True)
    X = otri.triangulate_poses(cams, obs, None, no_distortion=False, nviews=nviews)
    assert np.abs(X - poses).max() < 0.5


def test_undistort_matches_cv2_and_inverts_distort():
    cv2 = pytest.importorskip('cv2')
    cam = synth.camera_ring(4, seed=3)[1]
    rig = otri.build_multi_camera_system([('c', cam)], no_distortion=False)
    c = rig._cams['c']
    rng = np.random.default_rng(0)
    uv = rng.uniform(150, 850, (200, 2))
    K = np.array([[cam['fx'][0], 0, cam['cx'][0]], [0, cam['fy'][0], cam['cy'][0]], [0, 0, 1.0]])
    ref = cv2.undistortPoints(uv.reshape(-1, 1, 2), K, c.D, P=K).reshape(-1, 2)  # 5 iterations
    assert np.abs(c.undistort(uv) - ref).max() < 1e-8
    assert np.abs(c.distort(c.undistort(uv)) - uv).max() < 1e-3
    assert np.abs(c.M - K @ np.hstack([cam['R'], -cam['R'] @ cam['T']])).max() < 1e-6


def test_fewer_than_two_views_is_zero():
    poses, obs, cams = _rig_and_obs(4, 3, 1.0, distorted=True)
    vis = np.ones(obs.shape[:2])
    vis[0:3, 5] = 0                                  # frame 0, joint 5: one view left
    vis[4:8, 2] = 0                                  # frame 1, joint 2: none
    X = otri.triangulate_poses(cams, obs, vis)
    assert np.all(X[0, 5] == 0) and np.all(X[1, 2] == 0) and np.all(X[0, 4] != 0)
    proj, rv = otri.reproject_poses(obs, cams, vis)
    assert np.all(rv[0:4, 5] == 0) and np.all(proj[0:4, 5] == 0) and np.all(rv[0:4, 4] == 1)


def test_ransac_rejects_outlier_view_and_tie_order():
    poses, obs, cams = _rig_and_obs(4, 4, 0.3, distorted=True, seed=5)
    obs[2, 7] += 80.0                                # frame 0, view 2, joint 7 is an outlier
    vis = np.ones(obs.shape[:2])
    res = otri.ransac(obs, cams, vis, reproj_thre=10.0, num_inliers=3)
    assert res[0:4, 7].tolist() == [1, 1, 0, 1]
    assert np.all(res[4:8] == 1)
    # demanding 4 inliers drops the joint entirely
    res4 = otri.ransac(obs, cams, vis, reproj_thre=10.0, num_inliers=4)
    assert res4[0:4, 7].tolist() == [0, 0, 0, 0]
    # an originally invisible view can come back as an inlier (triangulate.py:164-165)
    vis2 = vis.copy()
    vis2[3, 1] = 0
    assert otri.ransac(obs, cams, vis2, 10.0, 3)[3, 1] == 1


# ---- epipolar -----------------------------------------------------------------------
def test_epipolar_zero_for_exact_geometry():
    rigs = synth.camera_table(2, 4, seed=7)
    poses = synth.random_poses(5, seed=8)
    subj = np.array([0, 1, 1, 0, 1])
    obs, _ = synth.multiview_observations(poses, rigs, subj, distorted=False)
    F = oepi.fundamental_table({s: rigs[s] for s in range(2)})
    r = oepi.epipolar_residuals(obs, subj, F)
    assert r.shape == (5, 12, 17) and r.max() < 1e-9
    obs2 = obs + np.random.default_rng(0).normal(0, 2, obs.shape)
    assert oepi.epipolar_residuals(obs2, subj, F).mean() > 1e-6
    views = [obs2.reshape(5, 4, 17, 2)[:, v] for v in range(4)]
    w = [np.ones((5, 17, 1)) for _ in range(4)]
    loss = oepi.fundamental_loss(views, w, subj, F, True)
    assert abs(loss - oepi.epipolar_residuals(obs2, subj, F).mean()) < 1e-12


# ---- RPSM -----------------------------------------------------------------------------
def test_body_trees():
    b = HumanBody()
    assert b.root_idx == 6 and len(b.edges()) == 15
    order = [n['idx'] for n in b.skeleton_sorted_by_level]
    seen = set()
    for j in order:                                   # children before parents
        assert all(c in seen for c in b.skeleton[j]['children'])
        seen.add(j)
    assert len(h36m17().edges()) == 16


def test_grid_layout():
    r = golden('rpsm.npz')
    assert np.array_equal(opict.compute_grid(125.0, np.array([1.0, 2.0, 3.0]), 2), r['grid2'])
    g = opict.compute_grid(2000, np.zeros(3), 16)
    lin = np.linspace(-1000, 1000, 16)
    b = 5 * 256 + 3 * 16 + 9                          # iy=5, ix=3, iz=9
    assert np.array_equal(g[b], [lin[3], lin[5], lin[9]])


def test_rpsm_matches_reference_golden():
    r = golden('rpsm.npz')
    body = HumanBody()
    cfg = rpsm_config()
    for f in range(2):
        hm, cams, boxes, root, limb, edges = rpsm_golden_frame(r, f)
        assert edges == body.edges()
        avg = {e: float(l) for e, l in zip(edges, r['avg_limb'])}
        if f == 0:
            pw = opict.level0_pairwise(2000, avg, 16, body)
        grid = opict.compute_grid(2000, root, 16)
        assert np.array_equal(grid, r['f%d_grid0' % f])
        unary = np.array(opict.compute_unary_term(hm, [grid], boxes, cams, cfg.NETWORK.IMAGE_SIZE))
        assert np.array_equal(unary, r['f%d_unary0' % f])
        pose, trace = opict.rpsm(cams, hm, boxes, root, limb, pw, cfg, body, return_trace=True)
        assert np.array_equal(trace, r['f%d_trace' % f])
        assert np.array_equal(pose, r['f%d_pose' % f])


def test_two_view_triangulation_matches_cv2_triangulatePoints():
    """Independent anchor for the restated pymvg find3d: OpenCV's own linear (DLT + SVD) two-view
    triangulation on the same projection matrices and the same (undistorted) observations."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(5)
    rig = synth.camera_ring(4, seed=9)
    poses = synth.random_poses(5, seed=10)
    obs, cams = synth.multiview_observations(poses, [rig], [0] * 5, noise_px=1.5, seed=11, distorted=False)
    system = otri.build_multi_camera_system([('camera_%d' % v, rig[v]) for v in range(4)], no_distortion=True)
    worst = 0.0
    for a, b in [(0, 1), (0, 3), (1, 2), (2, 3)]:
        Pa, Pb = system._cams['camera_%d' % a].M, system._cams['camera_%d' % b].M
        for f in range(5):
            xa, xb = obs[f * 4 + a].T.copy(), obs[f * 4 + b].T.copy()          # [2, J]
            Xh = cv2.triangulatePoints(Pa, Pb, xa, xb)                           # [4, J]
            ref = (Xh[:3] / Xh[3]).T
            for j in range(17):
                got = system.find3d([('camera_%d' % a, obs[f * 4 + a, j]), ('camera_%d' % b, obs[f * 4 + b, j])])
                worst = max(worst, np.linalg.norm(got - ref[j]))
    assert worst < 1e-6                                                        # mm


def test_affine_lu_emulation_vs_cv2_random():
    """The elimination order restated in oracle.transforms reproduces cv2.getAffineTransform bit for bit
    on thousands of random boxes (float32 and float64 inputs, with and without rotation)."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(123)
    bad = 0
    for i in range(1500):
        c = rng.uniform(0, 1200, 2)
        s = np.repeat(rng.uniform(0.2, 6.0), 2)
        if i % 2:
            c, s = c.astype(np.float32), s.astype(np.float32)
        rot = 0.0 if i % 3 else rng.uniform(-90, 90)
        size = [int(rng.choice([32, 48, 64, 80, 96, 256])), int(rng.choice([32, 48, 64, 80, 96, 256]))]
        src, dst = otr.affine_point_triples(c, s, rot, size)
        for frm, to in ((src, dst), (dst, src)):
            ref = cv2.getAffineTransform(np.float32(frm), np.float32(to))
            got = otr.solve_affine(frm, to)
            bad += int(not np.array_equal(ref, got))
    assert bad == 0


def test_bilinear_restatement_vs_scipy_rgi_random():
    """oracle.pictorial.bilinear_zero_outside against scipy's RegularGridInterpolator used exactly as
    lib/multiviews/pictorial.py:176-187 uses it: interior, edges, outside, NaN -- bit for bit."""
    interp = pytest.importorskip('scipy.interpolate')
    rng = np.random.default_rng(7)
    for h in (64, 17):
        hmap = rng.random((h, h)).astype(np.float32)
        xy = rng.uniform(-3, h + 2, (4000, 2))
        xy[:50] = np.round(xy[:50])                               # exactly on grid lines
        xy[50:60] = [h - 1, h - 1]
        xy[60:70] = [0.0, h - 1]
        xy[70:75] = [np.nextafter(h - 1.0, h), 3.0]               # just outside the upper edge
        xy[75:80] = [-1e-12, 5.0]
        xy[80:85] = [np.nan, 2.0]
        rgi = interp.RegularGridInterpolator(points=[np.arange(h), np.arange(h)], values=hmap.transpose(),
                                             bounds_error=False, fill_value=0)
        ref = rgi(xy)
        got = opict.bilinear_zero_outside(hmap, xy)
        assert np.array_equal(ref, got, equal_nan=True)


def _rvec_tvec(cam):
    """OpenCV extrinsics of an H36M camera dict: x_c = R X + t with t = -R T (triangulate.py:31-32)."""
    cv2 = pytest.importorskip('cv2')
    R = np.asarray(cam['R'], dtype=np.float64)
    rvec, _ = cv2.Rodrigues(R)
    tvec = -R.dot(np.asarray(cam['T'], dtype=np.float64).reshape(3, 1))
    K = np.array([[cam['fx'][0], 0, cam['cx'][0]], [0, cam['fy'][0], cam['cy'][0]], [0, 0, 1.0]])
    D = np.array([cam['k'][0, 0], cam['k'][1, 0], cam['p'][0, 0], cam['p'][1, 0], cam['k'][2, 0]])
    return rvec, tvec, K, D


def test_find2d_matches_cv2_projectPoints():
    """Independent anchor for the restated pymvg find2d (forward plumb-bob model): OpenCV's own
    projectPoints with the same K, distortion vector [k1,k2,p1,p2,k3] and extrinsics, with and without
    distortion, on points all over the capture volume."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(17)
    worst = 0.0
    for seed in range(4):
        rig = synth.camera_ring(4, seed=40 + seed)
        pts = rng.normal(0, 700, (200, 3)) + [0, 0, 900.0]
        for v, cam in enumerate(rig):
            rvec, tvec, K, D = _rvec_tvec(cam)
            for nd in (False, True):
                system = otri.build_multi_camera_system([('c', cam)], no_distortion=nd)
                ref, _ = cv2.projectPoints(pts.reshape(-1, 1, 3), rvec, tvec, K, np.zeros(5) if nd else D)
                got = np.array([system.find2d('c', p) for p in pts])
                worst = max(worst, np.abs(got - ref.reshape(-1, 2)).max())
    assert worst < 1e-8                                                        # px


@pytest.mark.parametrize('nviews', [3, 4, 8])
def test_n_view_find3d_matches_eigh_and_gesvd(nviews):
    """Independent anchors for the restated find3d beyond two views: the null vector of the stacked DLT
    rows from (a) numpy's symmetric eigen-solver on A^T A and (b) LAPACK's QR-iteration SVD (gesvd), a
    different algorithm from the divide-and-conquer gesdd behind np.linalg.svd."""
    import scipy.linalg
    rig = synth.camera_ring(nviews, seed=60 + nviews)
    poses = synth.random_poses(6, seed=61)
    obs, cams = synth.multiview_observations(poses, [rig], [0] * 6, noise_px=2.0, seed=62)
    system = otri.build_multi_camera_system([('camera_%d' % v, rig[v]) for v in range(nviews)])
    worst_eigh = worst_svd = 0.0
    for f in range(6):
        for j in range(17):
            pts = [('camera_%d' % v, obs[f * nviews + v, j]) for v in range(nviews)]
            got = system.find3d(pts)
            rows = []
            for name, xy in pts:
                cam = system._cams[name]
                x, y = cam.undistort(np.asarray(xy).reshape(1, 2))[0]
                rows += [x * cam.M[2] - cam.M[0], y * cam.M[2] - cam.M[1]]
            A = np.array(rows)
            w, vecs = np.linalg.eigh(A.T.dot(A))
            e = vecs[:, 0]
            worst_eigh = max(worst_eigh, np.linalg.norm(got - e[:3] / e[3]))
            _, _, vt = scipy.linalg.svd(A, lapack_driver='gesvd')
            worst_svd = max(worst_svd, np.linalg.norm(got - vt[-1, :3] / vt[-1, 3]))
    assert worst_svd < 1e-6 and worst_eigh < 1e-4                              # mm (eigh squares the condition number)


def test_load_camera_from_M_normalises_a_scaled_projection_matrix():
    """pymvg's load_camera_from_M divides by K[2,2] when the projection matrix arrives scaled
    (M -> a M): K, R, t, the projections and the triangulated points must not depend on the scale."""
    rng = np.random.default_rng(23)
    cam = synth.camera_ring(4, seed=70)[1]
    base = otri.build_multi_camera_system([('c', cam)])._cams['c']
    pts = rng.normal(0, 500, (20, 3)) + [0, 0, 900.0]
    from oracle.pymvg_restated import RestatedCamera, RestatedMultiCameraSystem
    for scale in (2.5, 0.01, -3.0, 1.0 + 1e-9):
        cam2 = RestatedCamera.load_camera_from_M(base.M * scale, name='c', distortion_coefficients=base.D)
        assert abs(cam2.K[2, 2] - 1.0) < 1e-12
        assert np.abs(cam2.K - base.K).max() < 1e-7 * np.abs(base.K).max()
        assert np.abs(cam2.R - base.R).max() < 1e-10
        assert np.abs(cam2.project_3d_to_pixel(pts) - base.project_3d_to_pixel(pts)).max() < 1e-7
    # and a two-camera rig built from scaled matrices triangulates to the same points
    rig = synth.camera_ring(4, seed=71)
    sys1 = otri.build_multi_camera_system([('camera_%d' % v, rig[v]) for v in range(4)])
    cams2 = [RestatedCamera.load_camera_from_M(sys1._cams['camera_%d' % v].M * (1.7 + v), name='camera_%d' % v,
                                               distortion_coefficients=sys1._cams['camera_%d' % v].D) for v in range(4)]
    sys2 = RestatedMultiCameraSystem(cams2)
    X = np.array([120.0, -80.0, 1000.0])
    obs = [('camera_%d' % v, sys1.find2d('camera_%d' % v, X)) for v in range(4)]
    assert np.linalg.norm(sys1.find3d(obs) - X) < 1e-6
    # the DLT rows scale with M, so the minimiser under noise does depend on the scale -- which is why the
    # product keeps the reference's K[2,2] = 1 construction (triangulate.py:29-36); noise-free it does not
    assert np.linalg.norm(sys2.find3d(obs) - X) < 1e-6
