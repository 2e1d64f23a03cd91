#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference code (build container only).

Imports the reference modules that import cleanly under numpy 2 from
``/root/reference/lib`` (core.inference, utils.transforms, multiviews.cameras,
multiviews.pictorial, multiviews.body -- SURVEY.md section 8c) and records their
outputs on seeded synthetic inputs.  ``/root/reference`` does not exist on the
GPU box, so the vectors are committed next to this script and every test reads
only the ``.npz`` files.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Nothing from the reference is copied: only numeric inputs/outputs are stored.
``multiviews.triangulate`` cannot be imported (needs pymvg, not installable
offline) -- there are no golden vectors for it; see oracle/pymvg_restated.py.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_LIB = '/root/reference/lib'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_LIB)

from core import inference as ref_inference            # noqa: E402
from utils import transforms as ref_transforms          # noqa: E402
from multiviews import cameras as ref_cameras           # noqa: E402
from multiviews import pictorial as ref_pictorial       # noqa: E402
from multiviews.body import HumanBody as RefHumanBody   # noqa: E402

from pose_unsupervised_b200.utils import synth          # noqa: E402
from oracle import pictorial as orc_pictorial           # noqa: E402
from oracle.body import HumanBody as OrcHumanBody       # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print('wrote', path, '%.1f KB' % (os.path.getsize(path) / 1024))


def golden_affine():
    rng = np.random.default_rng(11)
    n = 256
    center = rng.uniform(100, 900, (n, 2))
    scale = rng.uniform(0.5, 4.0, (n, 1)).repeat(2, 1)
    scale[n // 2:, 1] *= rng.uniform(0.8, 1.2, n - n // 2)      # scale[1] is ignored by the reference
    rot = np.zeros(n)
    rot[3 * n // 4:] = rng.uniform(-60, 60, n - 3 * n // 4)
    sizes = np.array([[64, 64], [96, 96], [80, 80], [48, 64], [256, 256]])
    size = sizes[rng.integers(0, len(sizes), n)]
    f32 = rng.random(n) < 0.5                                   # float32 center/scale as validate() feeds
    fwd = np.empty((n, 2, 3))
    inv = np.empty((n, 2, 3))
    for i in range(n):
        c = center[i].astype(np.float32) if f32[i] else center[i]
        s = scale[i].astype(np.float32) if f32[i] else scale[i]
        fwd[i] = ref_transforms.get_affine_transform(c, s, rot[i], size[i])
        inv[i] = ref_transforms.get_affine_transform(c, s, rot[i], size[i], inv=1)
    save('affine.npz', center=center, scale=scale, rot=rot, size=size, f32=f32, fwd=fwd, inv=inv)


def _decode_case(rng, n, j, h, w, kind):
    hm = rng.random((n, j, h, w), dtype=np.float32)
    if kind == 'ties':
        hm = np.round(hm * 8) / 8                               # plateaus: many exact maxima
    elif kind == 'negative':
        hm = hm - 2.0                                            # every max <= 0 -> coords masked to 0
    elif kind == 'border':
        flat = hm.reshape(n, j, -1)
        spots = [0, w - 1, w, 2 * w + 1, (h - 1) * w, h * w - 1, (h - 2) * w + w - 2, 2 * w + 2,
                 w + w // 2, (h - 2) * w + 2]
        for a in range(n):
            for b in range(j):
                flat[a, b, spots[(a * j + b) % len(spots)]] = 2.0
    elif kind == 'special':
        flat = hm.reshape(n, j, -1)
        flat[0, 0, :] = 0.0                                      # all zero -> idx 0, masked
        flat[0, 1, :] = -0.0
        flat[0, 1, 7] = 0.0                                      # -0 == +0: first index wins
        flat[0, 2, 100] = np.nan                                 # NaN counts as max, first NaN wins
        flat[0, 2, 200] = np.nan
        flat[0, 3, 5] = np.inf
        flat[0, 3, 9] = np.inf
        flat[1, 0, :] = 0.25                                     # constant plateau
        flat[1, 1, 300] = -np.float32(np.nan)                    # negative-signed NaN
        flat[1, 2, :] = -np.inf
        flat[1, 3, h * w - 1] = 5.0                              # last element
    return hm.astype(np.float32)


def golden_decode():
    rng = np.random.default_rng(5)
    out = {}
    cases = [('rand64', 8, 5, 64, 64, 'rand'), ('ties64', 4, 5, 64, 64, 'ties'),
             ('neg64', 2, 3, 64, 64, 'negative'), ('border64', 4, 5, 64, 64, 'border'),
             ('special64', 2, 4, 64, 64, 'special'), ('rand96', 2, 3, 96, 96, 'rand'),
             ('rect', 4, 3, 48, 32, 'rand'), ('odd', 3, 2, 17, 23, 'ties')]
    names = []
    for name, n, j, h, w, kind in cases:
        hm = _decode_case(rng, n, j, h, w, kind)
        center = rng.uniform(300, 700, (n, 2))
        scale = rng.uniform(1.0, 3.5, (n, 1)).repeat(2, 1)
        if name in ('rand64', 'rect'):
            center = center.astype(np.float32)
            scale = scale.astype(np.float32)
        preds, maxvals = ref_inference.get_max_preds(hm)
        idx = np.argmax(hm.reshape(n, j, -1), 2)
        out[name + '_hm'] = hm
        out[name + '_center'] = center
        out[name + '_scale'] = scale
        out[name + '_idx'] = idx
        out[name + '_preds'] = preds
        out[name + '_maxvals'] = maxvals
        for pp in (False, True):
            cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(POST_PROCESS=pp))
            with np.errstate(invalid='ignore'):
                fp, fm = ref_inference.get_final_preds(cfg, hm, center, scale)
            out[name + '_final%d' % pp] = fp
            assert np.array_equal(fm, maxvals, equal_nan=True)
        names.append(name)
    out['names'] = np.array(names)
    save('decode.npz', **out)


def golden_cameras():
    rng = np.random.default_rng(3)
    rigs = synth.camera_table(nsubjects=3, nviews=4, seed=2)
    pts = rng.normal(0, 700, (64, 3)) + np.array([0, 0, 900.0])
    proj, w2c, c2w, packed = [], [], [], []
    for rig in rigs:
        for cam in rig:
            proj.append(ref_cameras.project_pose(pts, cam))
            xc = ref_cameras.world_to_camera_frame(pts, cam['R'], cam['T'])
            w2c.append(xc)
            c2w.append(ref_cameras.camera_to_world_frame(xc, cam['R'], cam['T']))
            packed.append(np.concatenate([cam['R'].ravel(), cam['T'].ravel(), cam['fx'], cam['fy'],
                                          cam['cx'], cam['cy'], cam['k'].ravel(), cam['p'].ravel()]))
    save('cameras.npz', pts=pts, cams=np.array(packed), proj=np.array(proj),
         w2c=np.array(w2c), c2w=np.array(c2w))


def pack_cam(cam):
    return np.concatenate([cam['R'].ravel(), cam['T'].ravel(), cam['fx'], cam['fy'],
                           cam['cx'], cam['cy'], cam['k'].ravel(), cam['p'].ravel()])


def golden_rpsm():
    nframes = 2
    body = RefHumanBody()
    obody = OrcHumanBody()
    edges = obody.edges()
    cfg = types.SimpleNamespace(
        NETWORK=types.SimpleNamespace(IMAGE_SIZE=np.array([256, 256]), HEATMAP_SIZE=np.array([64, 64])),
        PICT_STRUCT=types.SimpleNamespace(FIRST_NBINS=16, RECUR_NBINS=2, RECUR_DEPTH=10,
                                          GRID_SIZE=2000, LIMB_LENGTH_TOLERANCE=150))
    poses = synth.random_poses(nframes, seed=21, njoints=16)
    avg_limb = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]])
                                  for p in synth.random_poses(64, seed=99, njoints=16)])) for e in edges}
    pairwise = orc_pictorial.level0_pairwise(2000, avg_limb, 16, obody)
    # spot-check the array predicate against the generator's literal formula
    # (run/test/generate_pairwise_constraints.py:88-93)
    grid0 = ref_pictorial.compute_grid(2000, np.zeros(3), 16)
    rng = np.random.default_rng(0)
    for e in edges[:3]:
        dense = pairwise[e].toarray()
        for i, j in rng.integers(0, 4096, (2000, 2)):
            lit = np.abs(np.linalg.norm(grid0[i] - grid0[j]) - avg_limb[e]) < 0.4 * avg_limb[e]
            assert bool(dense[i, j]) == bool(lit)
    out = {'edges': np.array(edges), 'avg_limb': np.array([avg_limb[e] for e in edges]),
           'poses_gt': poses}
    for f in range(nframes):
        cams = synth.camera_ring(4, seed=40 + f)
        boxes = synth.crop_box(cams, poses[f])
        hm = synth.gaussian_heatmaps(cams, boxes, poses[f], 64, 256, 2.0, 0.02, seed=f)
        q = np.clip(np.round(hm * 4096), 0, 65535).astype(np.uint16)   # 12-bit, exactly representable
        hm = q.astype(np.float32) / np.float32(4096)
        limb = synth.limb_lengths(poses[f], edges)
        root = poses[f][body.root_idx] + np.array([35.0, -20.0, 15.0])  # grid centre off the true root
        # stage-wise outputs of the real reference
        grid = ref_pictorial.compute_grid(2000, root, 16)
        unary = ref_pictorial.compute_unary_term(hm, [grid], boxes, cams, cfg.NETWORK.IMAGE_SIZE)
        idx0 = ref_pictorial.infer(unary, pairwise, body, cfg)
        pose = ref_pictorial.rpsm(cams, hm, boxes, root, limb, pairwise, cfg)
        # per-level trace through the reference's own functions
        trace = [np.array([b for _, b in idx0])]
        p3 = ref_pictorial.get_loc_from_cube_idx([grid], idx0)
        cur = 2000 / 16
        for _ in range(10):
            grids = [ref_pictorial.compute_grid(cur, p3[i], 2) for i in range(16)]
            un = ref_pictorial.compute_unary_term(hm, grids, boxes, cams, cfg.NETWORK.IMAGE_SIZE)
            pw = ref_pictorial.compute_pairwise_constrain(body.skeleton, limb, grids, 150)
            idx = ref_pictorial.infer(un, pw, body, cfg)
            p3 = ref_pictorial.get_loc_from_cube_idx(grids, idx)
            trace.append(np.array([b for _, b in idx]))
            cur = cur / 2
        assert np.array_equal(p3, pose)
        print('frame', f, 'rpsm MPJPE vs GT: %.2f mm' % np.mean(np.linalg.norm(pose - poses[f], axis=1)))
        out['f%d_hm_q12' % f] = q
        out['f%d_cams' % f] = np.array([pack_cam(c) for c in cams])
        out['f%d_box_center' % f] = np.array([b['center'] for b in boxes])
        out['f%d_box_scale' % f] = np.array([b['scale'] for b in boxes])
        out['f%d_root' % f] = root
        out['f%d_limb' % f] = np.array([limb[e] for e in edges])
        out['f%d_grid0' % f] = grid
        out['f%d_unary0' % f] = np.array(unary)
        out['f%d_trace' % f] = np.array(trace)
        out['f%d_pose' % f] = pose
    out['grid2'] = ref_pictorial.compute_grid(125.0, np.array([1.0, 2.0, 3.0]), 2)
    save('rpsm.npz', **out)


def main():
    golden_affine()
    golden_decode()
    golden_cameras()
    golden_rpsm()


if __name__ == '__main__':
    main()
