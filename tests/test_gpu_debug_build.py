"""The hot path through the debug-assert build of the library (libposeb200_debug.so, -DPB200_DEBUG_CHECKS=1).

compute-sanitizer is not offered on the GPU pool, so the kernels check their own hand-off invariants in a second
build: map indices and map counts of the decode schedule, candidate addresses / back pointers / task indices /
stage contents of the on-chip RPSM, item indices of RANSAC.  The script below runs in a subprocess with that
library loaded instead of the release one, shows that the counters are alive (a check that fails on purpose),
that none of the real checks fires, and that the results equal the release build's bit for bit."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEBUG_LIB = os.path.join(ROOT, 'pose_unsupervised_b200', 'libposeb200_debug.so')

SCRIPT = textwrap.dedent('''
    import ctypes, sys
    import numpy as np
    import torch
    sys.path.insert(0, %r)
    from pose_unsupervised_b200 import _lib, runtime as rt
    from pose_unsupervised_b200.multiviews import pictorial, triangulate
    from pose_unsupervised_b200.multiviews.body import HumanBody
    from pose_unsupervised_b200.utils import synth
    from tests.util import pseudo_config, rpsm_config
    lib = _lib.load()
    out = {}
    rng = np.random.default_rng(0)
    B, V, J = 1024, 4, 17
    rigs = synth.camera_table(3, V, seed=0)
    subj = rng.integers(0, 3, B)
    cams = [rigs[s][v] for s in subj for v in range(V)]
    g = torch.Generator(device='cuda').manual_seed(3)
    hm = torch.rand((B * V, J, 64, 64), generator=g, device='cuda')
    center = rng.uniform(400, 600, (B * V, 2)); scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    for dyn in (True, False):
        rt.set_decode_schedule(dyn)
        for _ in range(3):
            res = triangulate.lift_heatmaps(hm, center, scale, cams, conf_thre=0.3, return_idx=True)
        out['idx_%%d' %% dyn] = res.idx.cpu().numpy(); out['poses_%%d' %% dyn] = res.poses3d.cpu().numpy()
    rt.set_decode_schedule(True)
    poses = synth.random_poses(256, seed=1)
    obs, cams2 = synth.multiview_observations(poses, rigs, rng.integers(0, 3, 256), noise_px=2.0, outlier_frac=0.15, seed=2)
    vis = (rng.random(obs.shape[:2]) > 0.1).astype(np.float64)
    out['ransac'] = triangulate.ransac(obs, cams2, vis, pseudo_config(10.0, 3, False))
    body = HumanBody.h36m17(); edges = body.edges(); cfg = rpsm_config(depth=3)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)])) for e in edges}
    table = pictorial.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    frames = []
    for f in range(6):
        pose = synth.random_poses(1, seed=200 + f)[0]; rig = synth.camera_ring(4, seed=300 + f)
        boxes = synth.crop_box(rig, pose); h = synth.gaussian_heatmaps(rig, boxes, pose, 64, 256, 2.0, 0.02, seed=f)
        if f == 3: h[:, [0, 5, 9]] = -h[:, [0, 5, 9]] - 0.01
        if f == 4: h[:, [10, 15]] = 0.0
        frames.append((pose, rig, boxes, h, synth.limb_lengths(pose, edges)))
    pick = np.concatenate([np.arange(6), rng.integers(0, 6, 314)])
    p3, tr = pictorial.rpsm_batch([c for i in pick for c in frames[i][1]], np.array([frames[i][3] for i in pick]),
                                  np.array([b['center'] for i in pick for b in frames[i][2]]),
                                  np.array([b['scale'] for i in pick for b in frames[i][2]]),
                                  np.array([frames[i][0][0] for i in pick]), np.array([[frames[i][4][e] for e in edges] for i in pick]),
                                  table, cfg, body, return_trace=True)
    out['rpsm_pose'] = p3; out['rpsm_trace'] = tr
    counters = (ctypes.c_int32 * 16)()
    _lib.check(lib.pb200_debug_violations(counters, 1))
    out['counters'] = np.array(list(counters)); out['enabled'] = np.array(lib.pb200_debug_enabled())
    if lib.pb200_debug_enabled():
        _lib.check(lib.pb200_debug_selftest())
        _lib.check(lib.pb200_debug_violations(counters, 1))
        out['selftest'] = np.array(list(counters))
    np.savez(sys.argv[1], **out)
''') % ROOT


def _run(tmp_path, lib, name):
    path = str(tmp_path / (name + '.npz'))
    env = dict(os.environ)
    if lib:
        env['PB200_LIB'] = lib
    else:
        env.pop('PB200_LIB', None)
    res = subprocess.run([sys.executable, '-c', SCRIPT, path], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    return np.load(path)


def test_hot_path_through_the_debug_assert_build(tmp_path):
    if not os.path.exists(DEBUG_LIB):
        pytest.skip('libposeb200_debug.so is not built (python -m pose_unsupervised_b200.build --debug)')
    dbg = _run(tmp_path, DEBUG_LIB, 'debug')
    rel = _run(tmp_path, None, 'release')
    assert int(dbg['enabled']) == 1 and int(rel['enabled']) == 0
    assert dbg['selftest'][15] == 1 and dbg['selftest'][:15].sum() == 0          # the counters are alive ...
    assert not dbg['counters'].any(), dbg['counters']                            # ... and no real check fired
    for key in ('idx_1', 'idx_0', 'poses_1', 'poses_0', 'ransac', 'rpsm_pose', 'rpsm_trace'):
        assert np.array_equal(dbg[key], rel[key]), key
    assert np.array_equal(dbg['idx_1'], dbg['idx_0']) and np.array_equal(dbg['poses_1'], dbg['poses_0'])
