"""CPU test of the drop-in overlay: with the overlay directory in front of a
reference-style lib/ on sys.path, hot-path modules resolve to this repository and every other module of
the same packages still resolves to the reference's own file."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_overlay_resolution(tmp_path):
    lib = tmp_path / 'lib'
    for pkg, mod, body in [('core', 'function', 'from core.inference import get_final_preds\nWHO = "reference"\n'),
                           ('core', 'inference', 'WHO = "reference"\n'),
                           ('multiviews', 'other', 'WHO = "reference"\n'),
                           ('utils', 'vis', 'WHO = "reference"\n')]:
        d = lib / pkg
        d.mkdir(parents=True, exist_ok=True)
        (d / '__init__.py').write_text('')
        (d / (mod + '.py')).write_text(body)
    script = textwrap.dedent('''
        import sys
        sys.path.insert(0, %r)                                   # the reference's lib/
        sys.path.insert(0, %r)                                   # this repository
        sys.path.insert(0, %r)                                   # the overlay, in front
        import core.function, core.inference, multiviews.other, multiviews.triangulate, utils.vis, utils.transforms
        assert core.function.WHO == "reference" and multiviews.other.WHO == "reference" and utils.vis.WHO == "reference"
        assert "pose_unsupervised_b200" in core.inference.__file__
        assert core.function.get_final_preds.__module__ == "pose_unsupervised_b200.core.inference"
        assert multiviews.triangulate.triangulate_poses.__module__ == "pose_unsupervised_b200.multiviews.triangulate"
        assert utils.transforms.get_affine_transform.__module__ == "pose_unsupervised_b200.utils.transforms"
        print("overlay ok")
    ''') % (str(lib), ROOT, os.path.join(ROOT, 'pose_unsupervised_b200', 'dropin'))
    out = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'overlay ok' in out.stdout
