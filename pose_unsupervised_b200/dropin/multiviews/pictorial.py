"""multiviews.pictorial with the overlay in front (lib/multiviews/pictorial.py:19-250): ``rpsm``
is this repository's; the reference's building blocks (compute_grid, infer, ...) stay reachable."""
import multiviews as _pkg
from pose_unsupervised_b200.dropin._fallthrough import reference_names as _reference_names

_names, _reference = _reference_names(_pkg, 'pictorial', __file__)
globals().update(_names)

from pose_unsupervised_b200.multiviews.pictorial import (  # noqa: E402,F401
    rpsm, rpsm_batch, PairwiseTable, break_limb_length, lift_combination)
from pose_unsupervised_b200.multiviews.body import HumanBody  # noqa: E402,F401
