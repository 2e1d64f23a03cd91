"""Wire formats on either side of the lifting path ("next" row 3).

The reference hands data from the network side to the geometry side through HDF5 files:
``heatmaps_locations_<subset>_<dataset>.h5`` with datasets ``heatmaps [N,J,H,W]``,
``locations [N,J,3] = (x, y, maxval)`` and ``joint_names_order`` (lib/core/function.py:671-676), and
``<name>_pseudo_label.h5`` with ``pseudo_2d`` and ``joints_vis`` (run/test/test_pseudo_label.py:
213-216,255-258).  ``h5py`` is used when it is importable; otherwise the same dataset names are
stored in an ``.npz`` next to the requested path (``h5py`` is not in this image).  Host-side file
handling only -- no arithmetic of the lifting path lives here.
"""
import os

import numpy as np

try:
    import h5py
except ImportError:  # pragma: no cover - depends on the image
    h5py = None


def _npz_path(path):
    base, ext = os.path.splitext(str(path))
    return base + '.npz' if ext in ('.h5', '.hdf5') else str(path)


def write_datasets(path, **datasets):
    """Write named arrays to ``path`` (.h5 with h5py, else the .npz twin); returns the file written."""
    if h5py is not None and str(path).endswith(('.h5', '.hdf5')):
        with h5py.File(str(path), 'w') as f:
            for k, v in datasets.items():
                f[k] = np.asarray(v)
        return str(path)
    out = _npz_path(path)
    np.savez(out, **{k: np.asarray(v) for k, v in datasets.items()})
    return out


def read_datasets(path, names=None):
    """Read named arrays back (from the .h5 if it exists and h5py is present, else the .npz twin)."""
    if h5py is not None and os.path.exists(str(path)) and str(path).endswith(('.h5', '.hdf5')):
        with h5py.File(str(path), 'r') as f:
            return {k: np.array(f[k]) for k in (names or list(f.keys()))}
    with np.load(_npz_path(path)) as f:
        return {k: f[k] for k in (names or f.files)}


def write_heatmaps_locations(path, heatmaps, locations, joint_names_order):
    """lib/core/function.py:671-676."""
    return write_datasets(path, heatmaps=np.asarray(heatmaps, dtype=np.float32),
                          locations=np.asarray(locations, dtype=np.float32),
                          joint_names_order=np.asarray(joint_names_order))


def read_locations(path):
    """run/test/test_triangulate.py:60-63, test_pseudo_label.py:154-157 -> (pred2d [N,J,2], confidence [N,J])."""
    loc = read_datasets(path, ['locations'])['locations']
    return loc[:, :, :2], loc[:, :, 2]


def write_pseudo_label(path, pseudo_2d, joints_vis):
    """run/test/test_pseudo_label.py:213-216,255-258."""
    return write_datasets(path, pseudo_2d=pseudo_2d, joints_vis=joints_vis)


def read_pseudo_label(path):
    """lib/dataset/multiview_h36m_compatible.py:109-136 reads these two datasets."""
    d = read_datasets(path, ['pseudo_2d', 'joints_vis'])
    return d['pseudo_2d'], d['joints_vis']


def pareto_select(acc, num):
    """Indices kept by the Pareto selection over (PCKh, visible ratio) of
    run/test/test_pseudo_label.py:261-273 (same rank arithmetic, same tie behaviour)."""
    _, acc_order = np.unique(acc, return_inverse=True)
    _, num_order = np.unique(num, return_inverse=True)
    pending = list(np.argsort(acc_order + num_order))
    keep = []
    while pending:
        ref = pending.pop()
        keep.append(int(ref))
        pending = [i for i in pending
                   if not (acc_order[i] <= acc_order[ref] and num_order[i] <= num_order[ref])]
    return keep


def read_pairwise(path):
    """data/testdata/pairwise_b16.pkl (run/test/generate_pairwise_constraints.py:109-111):
    {'limb_length': {(p, c): mm}, 'pairwise_constrain': {(p, c): scipy sparse [n^3, n^3]}}."""
    import pickle
    with open(str(path), 'rb') as f:
        d = pickle.load(f)
    return d['limb_length'], d['pairwise_constrain']


def write_rpsm_testdata(path, records):
    """Write rows in the layout of run/test/generate_data_for_rpsm.py:110-117 (tests, examples)."""
    import pickle
    with open(str(path), 'wb') as f:
        pickle.dump(list(records), f)
    return str(path)


def read_rpsm_testdata(path, body, nviews=4):
    """data/testdata/rpsm_testdata_b16.pkl -> batched arrays for ``pictorial.rpsm_batch``.

    The pickle is a list of per-(frame, view) dicts ``{'heatmap' [J,h,w], 'cam_params', 'joints_3d_cam'
    [J,3], 'scale', 'center'}`` (run/test/generate_data_for_rpsm.py:110-117); grouping, ground truth,
    grid centre (root joint of view 0 in world coordinates) and per-frame limb lengths follow
    run/test/test_rpsm.py:81-126.  Returns a dict with ``cams`` (list of B*V camera dicts),
    ``heatmaps [B,V,J,h,w]``, ``centers [B*V,2]``, ``scales [B*V,2]``, ``grid_centers [B,3]``,
    ``limb_lengths [B,E]`` (``body.edges()`` order) and ``gt [B,J,3]``.
    """
    import pickle
    with open(str(path), 'rb') as f:
        db = pickle.load(f)
    assert len(db) % nviews == 0
    edges = body.edges()
    cams, hms, centers, scales, roots, limbs, gts = [], [], [], [], [], [], []
    for i in range(0, len(db), nviews):
        group = db[i:i + nviews]
        hms.append(np.array([g['heatmap'] for g in group], dtype=np.float32))
        cams += [g['cam_params'] for g in group]
        centers += [np.asarray(g['center'], dtype=np.float64).reshape(2) for g in group]
        scales += [np.asarray(g['scale'], dtype=np.float64).reshape(-1)[:2] for g in group]
        c0 = group[0]['cam_params']
        pose = (np.asarray(c0['R']).T.dot(np.asarray(group[0]['joints_3d_cam']).T) + np.asarray(c0['T']).reshape(3, 1)).T
        gts.append(pose)                                             # camera_to_world_frame of view 0
        roots.append(pose[body.root_idx])
        limbs.append([np.linalg.norm(pose[p] - pose[c]) for p, c in edges])
    return {'cams': cams, 'heatmaps': np.array(hms), 'centers': np.array(centers), 'scales': np.array(scales),
            'grid_centers': np.array(roots), 'limb_lengths': np.array(limbs), 'gt': np.array(gts)}
