"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header
declares, refuses to compute without a device, and the host-side logic around it."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from pose_unsupervised_b200 import _lib, parallel
from pose_unsupervised_b200.multiviews.body import HumanBody
from pose_unsupervised_b200.multiviews.cameras import pack_camera
from pose_unsupervised_b200.utils import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'poseb200.h')


@pytest.fixture(scope='module')
def lib():
    from pose_unsupervised_b200 import build
    build.build()                       # nvcc cross-compiles without a GPU
    return _lib.load()


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pb200_\w+)\s*\(', text)))


def test_header_and_binding_agree(lib):
    syms = header_symbols()
    assert len(syms) >= 18
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), s


def test_version_and_error_text(lib):
    assert lib.pb200_version() == 100
    if not torch.cuda.is_available():
        assert lib.pb200_device_check() != 0
        assert b'cuda' in lib.pb200_last_error().lower()


def test_argument_errors_need_no_device(lib):
    rc = lib.pb200_triangulate(None, None, None, 0, None, 1, 4, 17, 0, None, None)
    assert rc == -1 and b'null' in lib.pb200_last_error()
    one = ctypes.c_void_p(8)
    rc = lib.pb200_triangulate(one, one, one, 0, None, 1, 99, 17, 0, one, None)
    assert rc == -1 and b'V=99' in lib.pb200_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-device behaviour')
def test_no_cpu_fallback():
    from pose_unsupervised_b200.core.inference import get_max_preds
    from pose_unsupervised_b200.multiviews.triangulate import triangulate_poses
    with pytest.raises(_lib.Pb200Error):
        get_max_preds(np.zeros((1, 1, 8, 8), np.float32))
    cams = synth.camera_ring(4)
    with pytest.raises(_lib.Pb200Error):
        triangulate_poses(cams, np.zeros((4, 17, 2)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'pose_unsupervised_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_pack_camera_layout():
    cam = synth.camera_ring(4, seed=1)[2]
    rec = pack_camera(cam)
    assert rec.shape == (24,) and np.all(rec[21:] == 0)
    assert np.array_equal(rec[:9].reshape(3, 3), cam['R']) and np.array_equal(rec[9:12], cam['T'].ravel())
    assert rec[12] == cam['fx'][0] and rec[15] == cam['cy'][0]
    assert np.array_equal(rec[16:19], cam['k'].ravel()) and np.array_equal(rec[19:21], cam['p'].ravel())


def test_tree_arrays_match_reference_order():
    edges, order, root = HumanBody().tree_arrays()
    assert root == 6 and edges.shape == (15, 2) and edges.dtype == np.int32
    assert [tuple(e) for e in edges][:3] == [(1, 0), (2, 1), (3, 4)]
    pos = {int(j): i for i, j in enumerate(order)}
    assert all(pos[int(c)] < pos[int(p)] for p, c in edges)       # children first
    e17, o17, r17 = HumanBody.h36m17().tree_arrays()
    assert r17 == 0 and e17.shape == (16, 2) and sorted(o17.tolist()) == list(range(17))


def test_frame_shard_covers_everything():
    for n, w in [(10, 1), (10, 3), (4096, 8), (7, 8), (1000003, 8)]:
        spans = [parallel.frame_shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    rows = np.arange(24).reshape(12, 2)
    assert np.array_equal(parallel.shard_rows(rows, 4, 1, 3), rows[4:8])


def test_header_constants_match_binding():
    text = open(HEADER).read()
    consts = dict(re.findall(r'#define\s+(PB200_\w+)\s+\(?(-?\d+)\)?', text))
    assert int(consts['PB200_MAX_VIEWS']) == _lib.MAX_VIEWS
    assert int(consts['PB200_CAM_STRIDE']) == _lib.CAM_STRIDE
    assert int(consts['PB200_RPSM_MAX_JOINTS']) == _lib.RPSM_MAX_JOINTS
    assert (int(consts['PB200_F32']), int(consts['PB200_F64'])) == (_lib.F32, _lib.F64)
    assert int(consts['PB200_OK']) == 0 and int(consts['PB200_ERR_ARG']) == -1


def test_pairwise_bit_packing_layout():
    from pose_unsupervised_b200.multiviews.pictorial import pack_pairwise_bits
    rng = np.random.default_rng(0)
    for n in (27, 64, 100):
        m = (rng.random((n, n)) < 0.3).astype(np.int8)
        w = pack_pairwise_bits(m)
        assert w.dtype == np.uint32 and w.shape == (n, (n + 31) // 32)
        for i, j in rng.integers(0, n, (200, 2)):
            assert ((int(w[i, j // 32]) >> (j % 32)) & 1) == m[i, j]
    with pytest.raises(ValueError):
        pack_pairwise_bits(np.full((4, 4), 2))
    with pytest.raises(ValueError):
        pack_pairwise_bits(np.zeros((4, 5)))


def test_tuning_and_debug_entry_points_need_no_device(lib):
    """pb200_set_tuning validates key and value on the host; the release build reports that it is not a debug
    build, hands back zeroed counters and refuses the self-test."""
    import ctypes
    assert lib.pb200_set_tuning(_lib.TUNE_DECODE_SCHEDULE, _lib.DECODE_STATIC) == 0
    assert lib.pb200_set_tuning(_lib.TUNE_DECODE_SCHEDULE, _lib.DECODE_DYNAMIC) == 0
    assert lib.pb200_set_tuning(_lib.TUNE_DECODE_SCHEDULE, 7) != 0 and b'schedule' in lib.pb200_last_error()
    assert lib.pb200_set_tuning(99, 0) != 0 and b'unknown tuning key' in lib.pb200_last_error()
    assert lib.pb200_debug_enabled() == 0
    counters = (ctypes.c_int32 * 16)(*([5] * 16))
    assert lib.pb200_debug_violations(counters, 0) == 0 and not any(counters)
    assert lib.pb200_debug_selftest() != 0
    assert lib.pb200_debug_violations(None, 0) != 0


def test_debug_build_exports_the_same_abi():
    """libposeb200_debug.so (built by __graft_entry__.build()) exports every symbol of the header too."""
    import ctypes
    import os
    path = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libposeb200_debug.so')
    if not os.path.exists(path):
        pytest.skip('debug build not present')
    dbg = ctypes.CDLL(path)
    for name in _lib.EXPORTED_SYMBOLS:
        assert hasattr(dbg, name), name
    dbg.pb200_debug_enabled.restype = ctypes.c_int
    assert dbg.pb200_debug_enabled() == 1
