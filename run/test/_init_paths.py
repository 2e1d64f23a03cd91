"""Same role as the reference's run/test/_init_paths.py: put the B200 overlay of lib/ on sys.path
so that `from core.inference import ...` / `from multiviews.triangulate import ...` resolve to
libposeb200.  With the reference checkout present, add its lib/ AFTER these."""
import os.path as osp
import sys

this_dir = osp.dirname(osp.abspath(__file__))
repo = osp.abspath(osp.join(this_dir, '..', '..'))
for p in (repo, osp.join(repo, 'pose_unsupervised_b200', 'dropin')):
    if p not in sys.path:
        sys.path.insert(0, p)
