"""GPU parity of the RPSM kernel (K4) against the committed reference golden (real reference
run) and the oracle: identical chosen bins per level on generic inputs, identical final pose."""
import numpy as np
import pytest

from oracle import pictorial as opict
from oracle.body import HumanBody as OracleBody, h36m17
from pose_unsupervised_b200.utils import synth
from tests.util import golden, rpsm_config, rpsm_golden_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pict():
    from pose_unsupervised_b200.multiviews import pictorial
    return pictorial


def test_level0_pairwise_vs_oracle(pict):
    from pose_unsupervised_b200.multiviews.body import HumanBody
    r = golden('rpsm.npz')
    body, obody = HumanBody(), OracleBody()
    edges = obody.edges()
    avg = {e: float(l) for e, l in zip(edges, r['avg_limb'])}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    ref = opict.level0_pairwise(2000, avg, 16, obody, dense=True)
    for k in (0, 6, 14):
        assert np.array_equal(table.to_dense(k), ref[edges[k]])
    # the dict route (scipy sparse, as the reference's pickle holds them) gives the same bits
    sparse = opict.level0_pairwise(2000, avg, 16, obody)
    t2 = pict.PairwiseTable.from_dict(sparse, body)
    assert np.array_equal(t2.bits.cpu().numpy(), table.bits.cpu().numpy())


def test_rpsm_vs_reference_golden(pict):
    from pose_unsupervised_b200.multiviews.body import HumanBody
    r = golden('rpsm.npz')
    body, obody = HumanBody(), OracleBody()
    cfg = rpsm_config()
    avg = {e: float(l) for e, l in zip(obody.edges(), r['avg_limb'])}
    pw = opict.level0_pairwise(2000, avg, 16, obody)
    hms, cams, centers, scales, roots, limbs = [], [], [], [], [], []
    for f in range(2):
        hm, cam, boxes, root, limb, edges = rpsm_golden_frame(r, f)
        pose = pict.rpsm(cam, hm, boxes, root, limb, pw, cfg)         # reference signature, 1 frame
        assert pose.shape == (16, 3) and np.array_equal(pose, r['f%d_pose' % f]), f
        hms.append(hm); cams += cam; roots.append(root)
        centers += [b['center'] for b in boxes]; scales += [b['scale'] for b in boxes]
        limbs.append([limb[e] for e in edges])
    poses, trace = pict.rpsm_batch(cams, np.array(hms), np.array(centers), np.array(scales),
                                   np.array(roots), np.array(limbs), pw, cfg, body, return_trace=True)
    for f in range(2):
        assert np.array_equal(trace[f], r['f%d_trace' % f]), f
        assert np.array_equal(poses[f], r['f%d_pose' % f]), f


def test_rpsm_17_joints_vs_oracle_and_gt(pict):
    """BASELINE.json config 3: 4 views, 17 joints, 16^3 then 10 x 2^3."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    body, obody = HumanBody.h36m17(), h36m17()
    edges = obody.edges()
    cfg = rpsm_config()
    poses_gt = synth.random_poses(6, seed=31)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)]))
           for e in edges}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    opw = opict.level0_pairwise(2000, avg, 16, obody)
    hms, cams, centers, scales, roots, limbs = [], [], [], [], [], []
    for f in range(6):
        cam = synth.camera_ring(4, seed=50 + f)
        boxes = synth.crop_box(cam, poses_gt[f])
        hm = synth.gaussian_heatmaps(cam, boxes, poses_gt[f], 64, 256, 2.0, 0.02, seed=f)
        limb = synth.limb_lengths(poses_gt[f], edges)
        hms.append(hm); cams += cam; roots.append(poses_gt[f][0] + [20.0, -30.0, 10.0])
        centers += [b['center'] for b in boxes]; scales += [b['scale'] for b in boxes]
        limbs.append([limb[e] for e in edges])
    poses, trace = pict.rpsm_batch(cams, np.array(hms), np.array(centers), np.array(scales),
                                   np.array(roots), np.array(limbs), table, cfg, body, return_trace=True)
    for f in range(6):
        assert np.mean(np.linalg.norm(poses[f] - poses_gt[f], axis=1)) < 150.0
    for f in range(2):                                                   # oracle: ~2 s per frame
        boxes = [{'center': centers[f * 4 + v], 'scale': scales[f * 4 + v]} for v in range(4)]
        limb = {e: limbs[f][k] for k, e in enumerate(edges)}
        ref, rtrace = opict.rpsm(cams[f * 4:f * 4 + 4], hms[f], boxes, roots[f], limb, opw, cfg, obody,
                                 return_trace=True)
        # acceptance (hard part 4): identical bins on generic inputs,
        # else within one final-level cell (2000/16/2^10 = 0.12 mm)
        if not np.array_equal(trace[f], rtrace):
            assert np.abs(poses[f] - ref).max() < 0.13
        else:
            assert np.array_equal(poses[f], ref)


def test_rpsm_bit_rows_and_offset_table_agree(pict):
    """The shared-memory offset table and the bit-matrix rows are two readers of the same
    predicate; a matrix that is NOT a function of the offset must take the row path."""
    import torch
    from pose_unsupervised_b200.multiviews.body import HumanBody
    r = golden('rpsm.npz')
    body, obody = HumanBody(), OracleBody()
    cfg = rpsm_config()
    avg = {e: float(l) for e, l in zip(obody.edges(), r['avg_limb'])}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    assert table.offset_only
    hm, cam, boxes, root, limb, edges = rpsm_golden_frame(r, 0)
    args = (cam, hm[None], np.array([b['center'] for b in boxes]), np.array([b['scale'] for b in boxes]),
            root[None], np.array([[limb[e] for e in edges]]), table, cfg, body)
    p_lut, t_lut = pict.rpsm_batch(*args, return_trace=True, use_lut=True)
    p_row, t_row = pict.rpsm_batch(*args, return_trace=True, use_lut=False)
    assert np.array_equal(t_lut, t_row) and np.array_equal(p_lut, p_row)
    assert np.array_equal(t_lut[0], r['f0_trace'])
    # knock one bit out: no longer translation invariant -> detected -> row path, and the result
    # equals the oracle run on the same modified matrix
    bits = table.bits.clone()
    bits[6, 1234, 40] ^= 0x10
    broken = pict.PairwiseTable(bits, 4096)
    assert not broken.offset_only
    dense = {e: broken.to_dense(k) for k, e in enumerate(edges)}
    ref, rtrace = opict.rpsm(cam, hm, boxes, root, limb, dense, cfg, obody, return_trace=True)
    p_b, t_b = pict.rpsm_batch(*args[:6], broken, cfg, body, return_trace=True)
    assert np.array_equal(t_b[0], rtrace) and np.array_equal(p_b[0], ref)


def test_rpsm_zero_and_negative_energies(pict):
    """Disallowed children enter the reference's product as 0: when every allowed energy is <= 0
    the zero wins at the first disallowed index.  Negative / all-zero heatmaps exercise that."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    r = golden('rpsm.npz')
    body, obody = HumanBody(), OracleBody()
    cfg = rpsm_config(depth=3)
    avg = {e: float(l) for e, l in zip(obody.edges(), r['avg_limb'])}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    opw = opict.level0_pairwise(2000, avg, 16, obody)
    hm, cam, boxes, root, limb, edges = rpsm_golden_frame(r, 1)
    variants = []
    neg = hm.copy(); neg[:, [0, 5, 9]] = -neg[:, [0, 5, 9]] - 0.01          # three joints all negative
    variants.append(neg)
    zero = hm.copy(); zero[:, [10, 15]] = 0.0                                 # two joints all zero
    variants.append(zero)
    shifted = hm - 0.05                                                       # mixed signs everywhere
    variants.append(shifted.astype(np.float32))
    for h in variants:
        ref, rtrace = opict.rpsm(cam, h, boxes, root, limb, opw, cfg, obody, return_trace=True)
        got, trace = pict.rpsm_batch(cam, h[None], np.array([b['center'] for b in boxes]),
                                     np.array([b['scale'] for b in boxes]), root[None],
                                     np.array([[limb[e] for e in edges]]), table, cfg, body, return_trace=True)
        assert np.array_equal(trace[0], rtrace)
        assert np.array_equal(got[0], ref)


def _frame(body_edges, njoints, seed, nviews=4):
    poses = synth.random_poses(1, seed=seed, njoints=njoints)[0]
    cams = synth.camera_ring(nviews, seed=seed + 1)
    boxes = synth.crop_box(cams, poses)
    hm = synth.gaussian_heatmaps(cams, boxes, poses, 64, 256, 2.0, 0.02, seed=seed)
    limb = synth.limb_lengths(poses, body_edges)
    return poses, cams, boxes, hm, limb


def test_rpsm_generic_grid_sizes_vs_oracle(pict):
    """Non-default shapes take the generic code paths: 8^3 level-0 grid, 3^3 refinement grids
    (27 bins per joint, no shuffle merge), 2 views, depth 3."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    body, obody = HumanBody(), OracleBody()
    edges = obody.edges()
    cfg = rpsm_config(first=8, recur=3, depth=3)
    poses, cams, boxes, hm, limb = _frame(edges, 16, seed=71, nviews=2)
    avg = {e: limb[e] * 1.05 for e in edges}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 8)
    opw = opict.level0_pairwise(2000, avg, 8, obody)
    root = poses[6] + [15.0, -10.0, 5.0]
    ref, rtrace = opict.rpsm(cams, hm, boxes, root, limb, opw, cfg, obody, return_trace=True)
    got, trace = pict.rpsm_batch(cams, hm[None], np.array([b['center'] for b in boxes]),
                                 np.array([b['scale'] for b in boxes]), root[None],
                                 np.array([[limb[e] for e in edges]]), table, cfg, body, return_trace=True)
    assert np.array_equal(trace[0], rtrace) and np.array_equal(got[0], ref)


def test_rpsm_wide_shells_take_the_sorted_walk(pict):
    """Limbs longer than 5 grid cells exceed the enumeration reach: the offset-table path then sorts
    the child bins and walks to the first allowed one.  Same answer as the oracle."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    body, obody = HumanBody(), OracleBody()
    edges = obody.edges()
    cfg = rpsm_config(depth=2)
    poses, cams, boxes, hm, limb = _frame(edges, 16, seed=81)
    avg = {e: max(limb[e], 200.0) * 2.2 for e in edges}          # up to ~1000 mm: reach 7-10 cells of 133 mm
    assert max(avg.values()) * 1.4 / (2000 / 15) > 5.5
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    opw = opict.level0_pairwise(2000, avg, 16, obody)
    root = poses[6]
    ref, rtrace = opict.rpsm(cams, hm, boxes, root, limb, opw, cfg, obody, return_trace=True)
    for use_lut in (True, False):
        got, trace = pict.rpsm_batch(cams, hm[None], np.array([b['center'] for b in boxes]),
                                     np.array([b['scale'] for b in boxes]), root[None],
                                     np.array([[limb[e] for e in edges]]), table, cfg, body,
                                     return_trace=True, use_lut=use_lut)
        assert np.array_equal(trace[0], rtrace) and np.array_equal(got[0], ref), use_lut


@pytest.mark.parametrize('nviews,hw,nframes', [(4, 64, 330), (8, 64, 24), (4, 96, 24), (2, 80, 24), (3, 63, 12)])
def test_rpsm_onchip_equals_generic_kernel(pict, nviews, hw, nframes):
    """The on-chip level 0 (staged heatmaps, shared-memory energies, value-only max-product with two parents
    per lane, argmax resolved during back-tracking) and the generic kernel are two implementations of the same
    arithmetic: identical bins at every level and identical poses.  330 frames > 2 x 148 SMs: every persistent block processes several frames, so the
    cross-frame prefetch is exercised; 8 views / 96^2 / 80^2 maps need several staged groups per joint;
    63^2 maps are not 16-byte sized and are sampled with plain loads.  Degenerate frames (all-negative and
    all-zero joints) take the zero-energy shortcut of the forward pass and the all-nonpositive rule of the
    back-tracking in the middle of the batch."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    body, obody = HumanBody.h36m17(), h36m17()
    edges = obody.edges()
    cfg = rpsm_config(depth=4)
    cfg.NETWORK.HEATMAP_SIZE = np.array([hw, hw])
    base = 6
    rng = np.random.default_rng(nviews * 1000 + hw)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)]))
           for e in edges}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    assert table.offset_only and 0 < table.max_reach <= 5
    frames = []
    for f in range(base):
        pose = synth.random_poses(1, seed=200 + f)[0]
        cams = synth.camera_ring(nviews, seed=300 + f)
        boxes = synth.crop_box(cams, pose)
        hm = synth.gaussian_heatmaps(cams, boxes, pose, hw, 256, 2.0 * hw / 64, 0.02, seed=f)
        if f == 3:
            hm[:, [0, 5, 9]] = -hm[:, [0, 5, 9]] - 0.01
        if f == 4:
            hm[:, [10, 15]] = 0.0
        if f == 5:
            hm = (hm - 0.05).astype(np.float32)
        frames.append((pose, cams, boxes, hm, synth.limb_lengths(pose, edges)))
    pick = rng.integers(0, base, nframes)
    pick[:base] = np.arange(base)                       # every kind of frame is present
    hms = np.array([frames[i][3] for i in pick])
    cams = [c for i in pick for c in frames[i][1]]
    centers = np.array([b['center'] for i in pick for b in frames[i][2]])
    scales = np.array([b['scale'] for i in pick for b in frames[i][2]])
    roots = np.array([frames[i][0][0] for i in pick]) + rng.normal(0, 40.0, (nframes, 3))
    limbs = np.array([[frames[i][4][e] for e in edges] for i in pick])
    args = (cams, hms, centers, scales, roots, limbs, table, cfg, body)
    p_on, t_on = pict.rpsm_batch(*args, return_trace=True, onchip=True)
    p_gen, t_gen = pict.rpsm_batch(*args, return_trace=True, onchip=False)
    assert np.array_equal(t_on, t_gen)
    assert np.array_equal(p_on, p_gen)
    # and one frame of each kind against the oracle itself
    for kind in (0, 3, 4):
        f = int(np.where(pick == kind)[0][0])
        boxes = [{'center': centers[f * nviews + v], 'scale': scales[f * nviews + v]} for v in range(nviews)]
        limb = {e: limbs[f][k] for k, e in enumerate(edges)}
        ref, rtrace = opict.rpsm(cams[f * nviews:(f + 1) * nviews], hms[f], boxes, roots[f], limb,
                                 opict.level0_pairwise(2000, avg, 16, obody), cfg, obody, return_trace=True)
        assert np.array_equal(t_on[f], rtrace) and np.array_equal(p_on[f], ref), kind


def test_rpsm_onchip_spills_vectors_for_deep_trees(pict):
    """A caterpillar tree (every spine joint has a leaf as FIRST child) keeps one accumulator alive per
    spine joint: more live energy vectors than fit in shared memory, so some are spilled to scratch.
    Same bins as the generic kernel and as the oracle."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    J = 17
    children = [[] for _ in range(J)]
    spine = list(range(0, J, 2))                      # 0, 2, 4, ..., 16
    for a, b in zip(spine[:-1], spine[1:]):
        children[a] = [a + 1, b]                      # leaf first, then the rest of the spine
    names = ['j%d' % i for i in range(J)]
    body = HumanBody(names, children, 0)
    obody = OracleBody(names, children, 0)
    edges = obody.edges()
    cfg = rpsm_config(depth=2)
    pose = synth.random_poses(1, seed=5)[0]
    cams = synth.camera_ring(4, seed=6)
    boxes = synth.crop_box(cams, pose)
    hm = synth.gaussian_heatmaps(cams, boxes, pose, 64, 256, 2.0, 0.02, seed=7)
    limb = synth.limb_lengths(pose, edges)
    avg = {e: min(max(limb[e], 150.0), 420.0) for e in edges}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    assert table.offset_only and table.max_reach <= 5
    args = (cams, hm[None], np.array([b['center'] for b in boxes]), np.array([b['scale'] for b in boxes]),
            pose[0][None], np.array([[limb[e] for e in edges]]), table, cfg, body)
    p_on, t_on = pict.rpsm_batch(*args, return_trace=True, onchip=True)
    p_gen, t_gen = pict.rpsm_batch(*args, return_trace=True, onchip=False)
    assert np.array_equal(t_on, t_gen) and np.array_equal(p_on, p_gen)
    ref, rtrace = opict.rpsm(cams, hm, boxes, pose[0], limb, opict.level0_pairwise(2000, avg, 16, obody), cfg,
                             obody, return_trace=True)
    assert np.array_equal(t_on[0], rtrace) and np.array_equal(p_on[0], ref)


def test_rpsm_onchip_bushy_tree(pict):
    """A root with six children, each with a child and a grandchild chain of its own kind: more than four edges per
    tree depth (several four-edge steps per depth, in the level-0 back-tracking and in the refinement) and a root
    with more children than the refinement's schedule tables hold, so the 8-bin max-product takes its table-free
    form.  On-chip = generic kernel = oracle, bins at every level and poses."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    J = 17
    children = [[] for _ in range(J)]
    children[0] = [1, 2, 3, 4, 5, 6]
    for k in range(1, 7):
        children[k] = [k + 6]                          # 7..12
    children[7] = [13, 14]
    children[9] = [15]
    children[15] = [16]
    names = ['j%d' % i for i in range(J)]
    body = HumanBody(names, children, 0)
    obody = OracleBody(names, children, 0)
    edges = obody.edges()
    cfg = rpsm_config(depth=3)
    pose = synth.random_poses(1, seed=15)[0]
    cams = synth.camera_ring(4, seed=16)
    boxes = synth.crop_box(cams, pose)
    hm = synth.gaussian_heatmaps(cams, boxes, pose, 64, 256, 2.0, 0.02, seed=17)
    limb = synth.limb_lengths(pose, edges)
    avg = {e: min(max(limb[e], 150.0), 300.0) for e in edges}   # short limbs: the offset lists stay on chip
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    assert table.offset_only and table.max_reach <= 5
    nframes = 3
    args = (cams * nframes, np.repeat(hm[None], nframes, 0), np.array([b['center'] for b in boxes] * nframes),
            np.array([b['scale'] for b in boxes] * nframes), np.repeat(pose[0][None], nframes, 0),
            np.repeat(np.array([[limb[e] for e in edges]]), nframes, 0), table, cfg, body)
    p_on, t_on = pict.rpsm_batch(*args, return_trace=True, onchip=True)
    p_gen, t_gen = pict.rpsm_batch(*args, return_trace=True, onchip=False)
    assert np.array_equal(t_on, t_gen) and np.array_equal(p_on, p_gen)
    ref, rtrace = opict.rpsm(cams, hm, boxes, pose[0], limb, opict.level0_pairwise(2000, avg, 16, obody), cfg,
                             obody, return_trace=True)
    for f in range(nframes):
        assert np.array_equal(t_on[f], rtrace) and np.array_equal(p_on[f], ref)


def test_rpsm_onchip_long_limbs_overflow_the_lists(pict):
    """Limbs 20 % longer than H36M's: the child-offset lists of all 16 edges no longer fit next to the energy
    vectors (2 944 entries for 2 716 places), so the last edges go without and take the per-lane enumeration while
    the others keep the list walk -- in the same frame.  On-chip = generic kernel = oracle."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    body, obody = HumanBody.h36m17(), h36m17()
    edges = obody.edges()
    cfg = rpsm_config(depth=2)
    avg = {e: 1.2 * float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)]))
           for e in edges}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    assert table.offset_only and table.max_reach == 5
    pose, cams, boxes, hm, limb = _frame(edges, 17, 900)
    nframes = 3
    args = (cams * nframes, np.repeat(hm[None], nframes, 0), np.array([b['center'] for b in boxes] * nframes),
            np.array([b['scale'] for b in boxes] * nframes), np.repeat(pose[0][None], nframes, 0),
            np.repeat(np.array([[limb[e] for e in edges]]), nframes, 0), table, cfg, body)
    p_on, t_on = pict.rpsm_batch(*args, return_trace=True, onchip=True)
    p_gen, t_gen = pict.rpsm_batch(*args, return_trace=True, onchip=False)
    assert np.array_equal(t_on, t_gen) and np.array_equal(p_on, p_gen)
    ref, rtrace = opict.rpsm(cams, hm, boxes, pose[0], limb, opict.level0_pairwise(2000, avg, 16, obody), cfg,
                             obody, return_trace=True)
    for f in range(nframes):
        assert np.array_equal(t_on[f], rtrace) and np.array_equal(p_on[f], ref)


def test_rpsm_full_batch_is_repeatable(pict):
    """592 frames (4 per SM) three times: the on-chip kernel's stage hand-off, dynamic task hand-out and
    cross-frame prefetch must give identical bins and poses on every launch."""
    import torch
    from pose_unsupervised_b200.multiviews.body import HumanBody
    body, obody = HumanBody.h36m17(), h36m17()
    edges = obody.edges()
    cfg = rpsm_config()
    base = 8
    rng = np.random.default_rng(77)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)]))
           for e in edges}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    frames = []
    for f in range(base):
        pose = synth.random_poses(1, seed=500 + f)[0]
        cams = synth.camera_ring(4, seed=600 + f)
        boxes = synth.crop_box(cams, pose)
        frames.append((pose, cams, boxes, synth.gaussian_heatmaps(cams, boxes, pose, 64, 256, 2.0, 0.02, seed=f),
                       synth.limb_lengths(pose, edges)))
    pick = rng.integers(0, base, 592)
    hms = torch.from_numpy(np.array([frames[i][3] for i in range(base)])).cuda()[torch.from_numpy(pick).cuda()].contiguous()
    cams = [c for i in pick for c in frames[i][1]]
    centers = np.array([b['center'] for i in pick for b in frames[i][2]])
    scales = np.array([b['scale'] for i in pick for b in frames[i][2]])
    roots = np.array([frames[i][0][0] for i in pick]) + rng.normal(0, 40.0, (592, 3))
    limbs = np.array([[frames[i][4][e] for e in edges] for i in pick])
    first = None
    for _ in range(3):
        poses, trace = pict.rpsm_batch(cams, hms, centers, scales, roots, limbs, table, cfg, body, return_trace=True)
        if first is None:
            first = (poses.clone(), trace.clone())
        else:
            assert torch.equal(first[0], poses) and torch.equal(first[1], trace)


def test_rpsm_nonfinite_frames_do_not_leak(pict):
    """inf / NaN heatmaps switch THEIR frame to the per-lane enumeration (no shortcuts); what such a frame
    answers is deterministic but unspecified (the reference's own answer is a property of numpy's NaN ordering).
    The neighbours in the batch -- same block, same shared memory, before and after -- must answer exactly
    what they answer in a clean batch, and two launches must agree bit for bit."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    body, obody = HumanBody.h36m17(), h36m17()
    edges = obody.edges()
    cfg = rpsm_config(depth=3)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)]))
           for e in edges}
    table = pict.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    nframes, base = 320, 5                                # > 2 frames per persistent block
    frames = [_frame(edges, 17, 700 + f) for f in range(base)]
    pick = np.arange(nframes) % base
    hms = np.array([frames[i][3] for i in pick])
    cams = [c for i in pick for c in frames[i][1]]
    centers = np.array([b['center'] for i in pick for b in frames[i][2]])
    scales = np.array([b['scale'] for i in pick for b in frames[i][2]])
    roots = np.array([frames[i][0][0] for i in pick])
    limbs = np.array([[frames[i][4][e] for e in edges] for i in pick])
    clean, clean_t = pict.rpsm_batch(cams, hms, centers, scales, roots, limbs, table, cfg, body, return_trace=True)
    dirty = hms.copy()
    bad = np.arange(3, nframes, 7)
    for n, f in enumerate(bad):
        if n % 3 == 0:
            dirty[f, 0, 4, 10:20, 10:20] = np.inf         # overflow: inf, then inf * 0 = NaN further up the tree
        elif n % 3 == 1:
            dirty[f, 1, 9, 30, 30] = np.nan
        else:
            dirty[f, 2, 0] = -np.inf
    runs = [pict.rpsm_batch(cams, dirty, centers, scales, roots, limbs, table, cfg, body, return_trace=True)
            for _ in range(2)]
    good = np.setdiff1d(np.arange(nframes), bad)
    for poses, trace in runs:
        assert np.array_equal(np.asarray(poses)[good], np.asarray(clean)[good])
        assert np.array_equal(np.asarray(trace)[good], np.asarray(clean_t)[good])
        t = np.asarray(trace)[bad]
        assert t.min() >= 0 and t[:, 0].max() < 16 ** 3 and t[:, 1:].max() < 8   # bins stay inside their grids
    assert np.array_equal(np.asarray(runs[0][0]), np.asarray(runs[1][0]), equal_nan=True)
    assert np.array_equal(np.asarray(runs[0][1]), np.asarray(runs[1][1]))
