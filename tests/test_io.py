"""CPU tests of the wire formats and the Pareto selection (host-side logic, no GPU)."""
import numpy as np

from pose_unsupervised_b200.utils import io


def test_heatmaps_locations_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    hm = rng.random((8, 16, 8, 8), dtype=np.float32)
    loc = rng.random((8, 16, 3)).astype(np.float32)
    written = io.write_heatmaps_locations(tmp_path / 'heatmaps_locations_validation_multiview_h36m.h5', hm, loc,
                                          np.arange(16))
    assert written.endswith('.h5') or written.endswith('.npz')
    pred2d, conf = io.read_locations(tmp_path / 'heatmaps_locations_validation_multiview_h36m.h5')
    assert np.array_equal(pred2d, loc[:, :, :2]) and np.array_equal(conf, loc[:, :, 2])
    d = io.read_datasets(tmp_path / 'heatmaps_locations_validation_multiview_h36m.h5')
    assert np.array_equal(d['heatmaps'], hm) and np.array_equal(d['joint_names_order'], np.arange(16))


def test_pseudo_label_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    p2d, vis = rng.random((12, 16, 2)), (rng.random((12, 16)) > 0.5).astype(np.float64)
    io.write_pseudo_label(tmp_path / '0.7_1_pseudo_label.h5', p2d, vis)
    a, b = io.read_pseudo_label(tmp_path / '0.7_1_pseudo_label.h5')
    assert np.array_equal(a, p2d) and np.array_equal(b, vis)


def _pareto_reference(acc, num):
    # run/test/test_pseudo_label.py:261-273, verbatim control flow
    _, acc_order = np.unique(acc, return_inverse=True)
    _, num_order = np.unique(num, return_inverse=True)
    sum_order = list(np.argsort(acc_order + num_order))
    final = []
    while sum_order:
        ref_idx = sum_order.pop()
        final.append(ref_idx)
        remove = [r for r in sum_order if acc_order[r] <= acc_order[ref_idx] and num_order[r] <= num_order[ref_idx]]
        sum_order = [i for i in sum_order if i not in remove]
    return [int(i) for i in final]


def test_pareto_selection():
    acc = [0.904, 0.93, 0.95, 0.967, 0.91, 0.967]
    num = [0.95, 0.90, 0.85, 0.60, 0.99, 0.55]
    keep = io.pareto_select(acc, num)
    assert keep == _pareto_reference(acc, num)
    assert 5 not in keep and 3 in keep and 4 in keep          # (0.967, 0.55) is dominated by (0.967, 0.60)
    rng = np.random.default_rng(2)
    for _ in range(20):
        a, n = rng.random(9).round(2), rng.random(9).round(2)
        assert io.pareto_select(a, n) == _pareto_reference(a, n)


def test_rpsm_testdata_round_trip(tmp_path):
    """Rows in the reference's rpsm pickle layout -> batched arrays (run/test/test_rpsm.py:81-126)."""
    from pose_unsupervised_b200.multiviews.body import HumanBody
    from pose_unsupervised_b200.utils import synth
    body = HumanBody()
    poses = synth.random_poses(3, seed=1, njoints=16)
    records = []
    for f in range(3):
        cams = synth.camera_ring(4, seed=f)
        boxes = synth.crop_box(cams, poses[f])
        hm = synth.gaussian_heatmaps(cams, boxes, poses[f], 16, 256, 1.0, 0.0, seed=f)
        for v in range(4):
            cam_xyz = (cams[v]['R'] @ (poses[f].T - cams[v]['T'])).T
            records.append({'heatmap': hm[v], 'cam_params': cams[v], 'joints_3d_cam': cam_xyz,
                            'scale': boxes[v]['scale'], 'center': boxes[v]['center']})
    path = io.write_rpsm_testdata(tmp_path / 'rpsm_testdata_b16.pkl', records)
    d = io.read_rpsm_testdata(path, body)
    assert d['heatmaps'].shape == (3, 4, 16, 16, 16) and len(d['cams']) == 12
    assert np.abs(d['gt'] - poses).max() < 1e-9
    assert np.abs(d['grid_centers'] - poses[:, body.root_idx]).max() < 1e-9
    e0 = body.edges()[0]
    assert abs(d['limb_lengths'][1, 0] - np.linalg.norm(poses[1, e0[0]] - poses[1, e0[1]])) < 1e-9
