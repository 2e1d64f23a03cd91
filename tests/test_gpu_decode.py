"""GPU parity: heatmap decode (K1) through the reference-facing API (which calls the C ABI)
against the committed reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import inference as oinf
from oracle import transforms as otr
from tests.util import decode_config, golden, ulp_diff_f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def inf():
    from pose_unsupervised_b200.core import inference
    return inference


def test_golden_cases_bit_exact(inf):
    d = golden('decode.npz')
    for name in d['names']:
        hm, c, s = d[name + '_hm'], d[name + '_center'], d[name + '_scale']
        preds, maxvals = inf.get_max_preds(hm)
        assert preds.dtype == np.float32 and preds.shape == d[name + '_preds'].shape
        assert np.array_equal(preds, d[name + '_preds']), name
        assert maxvals.shape == d[name + '_maxvals'].shape
        assert np.array_equal(maxvals, d[name + '_maxvals'], equal_nan=True), name
        _, _, idx = inf.decode_heatmaps(hm, return_idx=True)
        assert np.array_equal(idx.cpu().numpy(), d[name + '_idx']), name
        for pp in (False, True):
            fp, fm = inf.get_final_preds(decode_config(pp), hm, c, s)
            ref = d[name + '_final%d' % pp]
            assert np.array_equal(np.isnan(fp), np.isnan(ref)), (name, pp)
            assert ulp_diff_f32(fp, ref).max() <= 1, (name, pp)
            assert (ulp_diff_f32(fp, ref) > 0).mean() < 1e-3, (name, pp)
            assert np.array_equal(fm, d[name + '_maxvals'], equal_nan=True)


def test_crop_affine_bit_exact(inf):
    from pose_unsupervised_b200.utils.transforms import crop_affine, get_affine_transform
    a = golden('affine.npz')
    rot0 = np.where(a['rot'] == 0)[0]
    for dt in (np.float32, np.float64):
        sel = rot0[a['f32'][rot0] == (dt == np.float32)]
        for size in np.unique(a['size'][sel], axis=0):
            rows = sel[np.all(a['size'][sel] == size, axis=1)]
            c, s = a['center'][rows].astype(dt), a['scale'][rows].astype(dt)
            fwd = crop_affine(c, s, size, inv=0).cpu().numpy()
            inv = crop_affine(c, s, size, inv=1).cpu().numpy()
            assert np.array_equal(fwd, a['fwd'][rows]) and np.array_equal(inv, a['inv'][rows])
    i = int(rot0[0])
    c, s = a['center'][i], a['scale'][i]
    if a['f32'][i]:
        c, s = c.astype(np.float32), s.astype(np.float32)
    assert np.array_equal(get_affine_transform(c, s, 0, a['size'][i], inv=1), a['inv'][i])


@pytest.mark.parametrize('shape', [(64, 17, 64, 64), (12, 17, 96, 96), (6, 5, 80, 80), (5, 3, 33, 47)])
def test_random_vs_oracle(inf, shape):
    rng = np.random.default_rng(shape[0])
    hm = rng.random(shape, dtype=np.float32)
    hm[::3] = np.round(hm[::3] * 16) / 16                       # plateaus -> ties
    n = shape[0]
    c = rng.uniform(300, 700, (n, 2))
    s = np.repeat(rng.uniform(1.5, 3.0, (n, 1)), 2, axis=1)
    _, _, idx = inf.decode_heatmaps(hm, return_idx=True)
    assert np.array_equal(idx.cpu().numpy(), oinf.flat_argmax(hm))
    for pp in (False, True):
        fp, fm = inf.get_final_preds(decode_config(pp), hm, c, s)
        rp, rm = oinf.get_final_preds(pp, hm, c, s)
        assert np.array_equal(fm, rm)
        assert ulp_diff_f32(fp, rp).max() <= 1
        assert (ulp_diff_f32(fp, rp) > 0).mean() < 1e-3


def test_view_list_input_interleaves_rows(inf):
    """validate() holds V tensors [B,J,H,W]; decoded rows must come out frame*V + view."""
    rng = np.random.default_rng(1)
    B, V, J = 5, 4, 7
    views = [rng.random((B, J, 64, 64), dtype=np.float32) for _ in range(V)]
    stacked = np.stack(views, axis=1).reshape(B * V, J, 64, 64)
    xy_l, mv_l, idx_l = inf.decode_heatmaps([torch.from_numpy(v).cuda() for v in views], return_idx=True)
    xy_s, mv_s, idx_s = inf.decode_heatmaps(stacked, return_idx=True)
    assert torch.equal(xy_l, xy_s) and torch.equal(mv_l, mv_s) and torch.equal(idx_l, idx_s)


def test_unaligned_view_falls_back_to_exact_scan(inf):
    rng = np.random.default_rng(2)
    buf = torch.from_numpy(rng.random(3 * 4 * 64 * 64 + 1, dtype=np.float32)).cuda()
    hm = buf[1:].view(3, 4, 64, 64)                               # 4-byte aligned only
    assert hm.data_ptr() % 16 != 0
    _, _, idx = inf.decode_heatmaps(hm, return_idx=True)
    assert np.array_equal(idx.cpu().numpy(), oinf.flat_argmax(hm.cpu().numpy()))


def test_empty_and_bad_inputs(inf):
    xy, mv = inf.decode_heatmaps(np.zeros((0, 17, 64, 64), np.float32))
    assert xy.shape == (0, 17, 2) and mv.shape == (0, 17)
    with pytest.raises(TypeError):
        inf.decode_heatmaps(np.zeros((1, 1, 8, 8), np.float64))
    with pytest.raises(AssertionError):
        inf.get_max_preds(np.zeros((1, 8, 8), np.float32))


def test_full_size_properties(inf):
    """BASELINE.json config 2 size (4096 frames x 4 views x 17 joints x 64x64): size-independent
    checks on the device + oracle on a slice."""
    g = torch.Generator(device='cuda').manual_seed(0)
    B, V, J = 4096, 4, 17
    hm = torch.rand((B * V, J, 64, 64), generator=g, device='cuda', dtype=torch.float32)
    xy, mv, idx = inf.decode_heatmaps(hm, return_idx=True)
    flat = hm.view(B * V, J, -1)
    assert torch.equal(mv, flat.amax(dim=2))                               # the maximum
    assert torch.equal(torch.gather(flat, 2, idx.long()[..., None])[..., 0], mv)   # attained at idx
    pos = torch.arange(4096, device='cuda')
    for lo in range(0, B * V, 2048):                                        # first index attaining it
        eq = flat[lo:lo + 2048] == mv[lo:lo + 2048, :, None]
        first = torch.where(eq, pos, 4096).amin(dim=2)
        assert torch.equal(first.int(), idx[lo:lo + 2048])
    assert torch.equal(xy[..., 0], (idx % 64).float()) and torch.equal(xy[..., 1], (idx // 64).float())
    sl = slice(5000, 5064)
    assert np.array_equal(idx[sl].cpu().numpy(), oinf.flat_argmax(hm[sl].cpu().numpy()))


@pytest.mark.parametrize('shift', [True, False])
def test_flip_test_fusion_matches_torch_recipe(inf, shift):
    """validate()'s flip test (lib/core/function.py:567-583) restated with the reference's own torch ops."""
    rng = np.random.default_rng(3)
    B, V, J, hw = 6, 4, 16, 64
    flip_pairs = [[0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13]]
    out = [torch.from_numpy(rng.random((B, J, hw, hw), dtype=np.float32)).cuda() for _ in range(V)]
    out_f = [torch.from_numpy(rng.random((B, J, hw, hw), dtype=np.float32)).cuda() for _ in range(V)]
    center = rng.uniform(400, 600, (B * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    # reference recipe
    order = list(range(J))
    for a, b in flip_pairs:
        order[a], order[b] = b, a
    fb = [torch.index_select(torch.flip(v, dims=[3]), 1, torch.tensor(order).cuda()) for v in out_f]
    if shift:
        for v in fb:
            v[:, :, :, 1:] = v.clone()[:, :, :, 0:-1]
    ref = [(a + b) * 0.5 for a, b in zip(out, fb)]
    ref_rows = torch.stack(ref, dim=1).reshape(B * V, J, hw, hw)            # preds[k::nviews] layout
    avg, xy, mv, idx = inf.decode_heatmaps_flip(out, out_f, flip_pairs, shift, center, scale, True, return_idx=True)
    assert torch.equal(avg, ref_rows)
    xy2, mv2, idx2 = inf.decode_heatmaps(ref_rows, center, scale, post_process=True, return_idx=True)
    assert torch.equal(idx, idx2) and torch.equal(mv, mv2) and torch.equal(xy, xy2)


def test_ties_across_lanes_and_chunks(inf):
    """Equal maxima planted at two positions of the map: the smaller flat index must win whichever
    lanes, float4 slots or ring chunks the two positions fall into."""
    rng = np.random.default_rng(11)
    n, hw = 256, 64
    hm = rng.random((n, 1, hw, hw), dtype=np.float32) * 0.5
    flat = hm.reshape(n, -1)
    special = [0, 1, 3, 4, 127, 128, 1023, 1024, 1025, 2047, 2048, 3071, 3072, 4094, 4095]
    first = np.empty(n, dtype=np.int64)
    for i in range(n):
        a, b = (rng.choice(special, 2, replace=False) if i % 2 else rng.choice(hw * hw, 2, replace=False))
        flat[i, a] = flat[i, b] = 0.75
        if i % 5 == 0:                                  # a third, later copy
            c = max(a, b) + (hw * hw - 1 - max(a, b)) // 2
            flat[i, c] = 0.75
        first[i] = min(a, b)
    _, mv, idx = inf.decode_heatmaps(hm, return_idx=True)
    assert np.array_equal(idx.cpu().numpy()[:, 0], first)
    assert np.all(mv.cpu().numpy() == np.float32(0.75))
    view = torch.from_numpy(hm).cuda()
    buf = torch.empty(hm.size + 1, dtype=torch.float32, device='cuda')
    buf[1:] = view.reshape(-1)
    _, _, idx2 = inf.decode_heatmaps(buf[1:].view(n, 1, hw, hw), return_idx=True)   # unaligned: LDG path
    assert np.array_equal(idx2.cpu().numpy()[:, 0], first)
