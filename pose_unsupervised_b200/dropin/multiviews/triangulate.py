"""multiviews.triangulate with the overlay in front (lib/multiviews/triangulate.py:57-213).
The reference's file needs ``pymvg``; where that is absent nothing falls through and the module
holds this repository's functions only."""
import multiviews as _pkg
from pose_unsupervised_b200.dropin._fallthrough import reference_names as _reference_names

_names, _reference = _reference_names(_pkg, 'triangulate', __file__)
globals().update(_names)

from pose_unsupervised_b200.multiviews.triangulate import (  # noqa: E402,F401
    triangulate_poses, ransac, reproject_poses, lift_heatmaps, mpjpe_stats, mpjpe_summary)
