"""CPU tests of the kernels' float64 arithmetic.

csrc/lift_math.cuh is plain C++ marked __host__ __device__; tests/hostcheck builds it
with g++ (TEST-ONLY, never loaded by the package) so that the exact device arithmetic
can be compared with the oracle and the reference goldens without a GPU.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import triangulate as otri
from pose_unsupervised_b200.multiviews.cameras import pack_camera
from pose_unsupervised_b200.utils import synth
from tests.util import golden

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostcheck', 'hostcheck.cpp')
SO = os.path.join(HERE, 'hostcheck', '_hostcheck.so')


@pytest.fixture(scope='module')
def hc():
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    hdr = os.path.join(HERE, '..', 'pose_unsupervised_b200', 'csrc', 'lift_math.cuh')
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-x', 'c++',
                               '-o', SO, SRC])
    return ctypes.CDLL(SO)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _hc_affine(hc, c, s, rot, size, inv, shift=None):
    out = np.zeros(6)
    sincos = None
    if rot != 0:
        r = np.pi * rot / 180
        sincos = np.array([np.sin(r), np.cos(r)])
    sh = np.zeros(2, np.float32) if shift is None else shift
    hc.hc_crop_affine.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    hc.hc_crop_affine(P(c), int(c.dtype == np.float64), P(s), int(s.dtype == np.float64),
                      P(sincos) if sincos is not None else None, float(sh[0]), float(sh[1]),
                      int(sh.dtype == np.float64), 1, int(size[0]), int(size[1]), inv, P(out))
    return out


def test_crop_affine_bit_exact(hc):
    """Every golden matrix of the real get_affine_transform (cv2), rotated boxes included."""
    a = golden('affine.npz')
    assert (a['rot'] != 0).sum() >= 32
    for i in range(len(a['rot'])):
        c = np.ascontiguousarray(a['center'][i:i + 1])
        s = np.ascontiguousarray(a['scale'][i:i + 1])
        if a['f32'][i]:
            c, s = c.astype(np.float32), s.astype(np.float32)
        for inv in (0, 1):
            out = _hc_affine(hc, c, s, float(a['rot'][i]), a['size'][i], inv)
            assert np.array_equal(out, a['inv' if inv else 'fwd'][i].ravel()), (i, inv)


def test_crop_affine_shift_matches_oracle(hc):
    """`shift` != 0 (never used by the reference's callers, but part of the signature): every
    dtype combination against the oracle restatement, which is itself pinned by affine.npz."""
    from oracle import transforms as otr
    rng = np.random.default_rng(5)
    for trial in range(64):
        cd = np.float32 if trial & 1 else np.float64
        sd = np.float32 if trial & 2 else np.float64
        hd = np.float32 if trial & 4 else np.float64
        c = rng.uniform(100, 900, (1, 2)).astype(cd)
        s = np.repeat(rng.uniform(0.5, 4.0, (1, 1)), 2, 1).astype(sd)
        sh = rng.uniform(-0.3, 0.3, 2).astype(hd)
        rot = float(rng.uniform(-45, 45)) if trial & 8 else 0.0
        for inv in (0, 1):
            ref = otr.get_affine_transform(c[0], s[0], rot, [64, 48], shift=sh, inv=inv)
            out = _hc_affine(hc, c, s, rot, (64, 48), inv, shift=sh)
            assert np.array_equal(out, ref.ravel()), (trial, inv)


def test_project_h36m_vs_reference(hc):
    c = golden('cameras.npz')
    pts = np.ascontiguousarray(c['pts'])
    for i, v in enumerate(c['cams']):
        pk = np.zeros(24)
        pk[:21] = v
        out = np.zeros((len(pts), 2))
        hc.hc_project(P(pk), P(pts), len(pts), 0, P(out))
        # bit-exact on hosts whose BLAS accumulates R(x-T) with FMA; 1e-12 px otherwise
        assert np.abs(out - c['proj'][i]).max() < 1e-10


def test_unary_and_grid_vs_reference(hc):
    r = golden('rpsm.npz')
    for f in range(2):
        hm = r['f%d_hm_q12' % f].astype(np.float32) / np.float32(4096)
        grid = np.ascontiguousarray(r['f%d_grid0' % f])
        g2 = np.zeros_like(grid)
        hc.hc_grid(ctypes.c_double(2000.0), 16, P(np.ascontiguousarray(r['f%d_root' % f])), P(g2))
        assert np.array_equal(g2, grid)
        un = np.zeros((16, 4096))
        for v in range(4):
            pk = np.zeros(24)
            pk[:21] = r['f%d_cams' % f][v]
            aff = np.zeros(6)
            cc = np.ascontiguousarray(r['f%d_box_center' % f][v:v + 1])
            ss = np.ascontiguousarray(r['f%d_box_scale' % f][v:v + 1])
            aff[:] = _hc_affine(hc, cc, ss, 0.0, (256, 256), 0)
            for j in range(16):
                o = np.zeros(4096)
                h = np.ascontiguousarray(hm[v, j])
                hc.hc_unary(P(h), 64, 64, P(pk), P(aff), P(grid), 4096, ctypes.c_double(256.0),
                            ctypes.c_double(256.0), P(o))
                un[j] = un[j] + o
        gu = r['f%d_unary0' % f]
        assert np.abs(un - gu).max() <= 1e-12 * max(1.0, np.abs(gu).max())


@pytest.mark.parametrize('nd', [0, 1])
@pytest.mark.parametrize('noise', [0.0, 2.0, 30.0])
def test_jacobi_dlt_vs_svd_oracle(hc, nd, noise):
    """float64 Gram + Jacobi eigenvector against the oracle's SVD: budget 1e-2 mm (north_star)."""
    rng = np.random.default_rng(0)
    rigs = synth.camera_table(3, 4, seed=1)
    poses = synth.random_poses(12, seed=2)
    obs, cams = synth.multiview_observations(poses, rigs, rng.integers(0, 3, 12), noise_px=noise, seed=3,
                                             distorted=not nd)
    ref = otri.triangulate_poses(cams, obs, None, bool(nd))
    worst = 0.0
    for i in range(12):
        pk = np.array([pack_camera(cams[i * 4 + v]) for v in range(4)])
        for j in range(17):
            xy = np.ascontiguousarray(obs[i * 4:(i + 1) * 4, j])
            X = np.zeros(3)
            assert hc.hc_triangulate(P(pk), P(xy), None, 4, nd, P(X)) == 4
            worst = max(worst, np.linalg.norm(X - ref[i, j]))
    assert worst < 1e-6


def test_two_view_pairs_vs_svd_oracle(hc):
    """RANSAC triangulates from pairs: the worst-conditioned use of the Gram matrix."""
    rng = np.random.default_rng(1)
    rigs = synth.camera_table(2, 4, seed=4)
    poses = synth.random_poses(8, seed=5)
    obs, cams = synth.multiview_observations(poses, rigs, rng.integers(0, 2, 8), noise_px=3.0, seed=6)
    worst = 0.0
    for i in range(8):
        pk = np.array([pack_camera(cams[i * 4 + v]) for v in range(4)])
        for a, b in [(0, 1), (0, 2), (1, 3), (2, 3)]:
            vis = np.zeros((32, 17))
            vis[i * 4 + a] = vis[i * 4 + b] = 1
            ref = otri.triangulate_poses(cams, obs, vis)[i]
            m = np.zeros(4, np.uint8)
            m[a] = m[b] = 1
            for j in range(0, 17, 4):
                xy = np.ascontiguousarray(obs[i * 4:(i + 1) * 4, j])
                X = np.zeros(3)
                hc.hc_triangulate(P(pk), P(xy), P(m), 4, 0, P(X))
                worst = max(worst, np.linalg.norm(X - ref[j]))
    assert worst < 1e-5


def test_plumb_bob_and_undistort_vs_oracle(hc):
    """Device forward model (pymvg find2d) and the 5-iteration undistortion against the oracle's
    restatement, including the no-distortion mode."""
    rng = np.random.default_rng(3)
    for seed in range(3):
        cam = synth.camera_ring(4, seed=seed)[seed % 4]
        rig = otri.build_multi_camera_system([('c', cam)], no_distortion=False)
        rig_nd = otri.build_multi_camera_system([('c', cam)], no_distortion=True)
        pk = pack_camera(cam)
        pts = np.ascontiguousarray(rng.normal(0, 600, (50, 3)) + [0, 0, 900.0])
        for model, system, distorted in ((1, rig, True), (2, rig_nd, True)):
            out = np.zeros((50, 2))
            hc.hc_project(P(pk), P(pts), 50, model, P(out))
            ref = np.array([system.find2d('c', p, distorted=distorted) for p in pts])
            assert np.abs(out - ref).max() < 1e-8
        uv = np.ascontiguousarray(rng.uniform(100, 900, (50, 2)))
        for nd, system in ((0, rig), (1, rig_nd)):
            out = np.zeros((50, 2))
            hc.hc_undistort(P(pk), nd, P(uv), 50, P(out))
            assert np.abs(out - system._cams['c'].undistort(uv)).max() < 1e-9


def test_device_plumb_bob_vs_cv2_projectPoints(hc):
    """The kernels' forward model (csrc/lift_math.cuh::project_plumb_bob, compiled for the host) against
    OpenCV's own projectPoints: an anchor that does not go through the oracle's restatement at all."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(31)
    worst = 0.0
    for seed in range(4):
        rig = synth.camera_ring(4, seed=80 + seed)
        pts = np.ascontiguousarray(rng.normal(0, 700, (100, 3)) + [0, 0, 900.0])
        for cam in rig:
            R = np.asarray(cam['R'], dtype=np.float64)
            rvec, _ = cv2.Rodrigues(R)
            tvec = -R.dot(np.asarray(cam['T'], dtype=np.float64).reshape(3, 1))
            K = np.array([[cam['fx'][0], 0, cam['cx'][0]], [0, cam['fy'][0], cam['cy'][0]], [0, 0, 1.0]])
            D = np.array([cam['k'][0, 0], cam['k'][1, 0], cam['p'][0, 0], cam['p'][1, 0], cam['k'][2, 0]])
            pk = pack_camera(cam)
            for model, dist in ((1, D), (2, np.zeros(5))):
                out = np.zeros((100, 2))
                hc.hc_project(P(pk), P(pts), 100, model, P(out))
                ref, _ = cv2.projectPoints(pts.reshape(-1, 1, 3), rvec, tvec, K, dist)
                worst = max(worst, np.abs(out - ref.reshape(-1, 2)).max())
    assert worst < 1e-8
