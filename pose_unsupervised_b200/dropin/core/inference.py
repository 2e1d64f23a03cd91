"""core.inference with the overlay in front (lib/core/inference.py:19-75)."""
import core as _pkg
from pose_unsupervised_b200.dropin._fallthrough import reference_names as _reference_names

_names, _reference = _reference_names(_pkg, 'inference', __file__)
globals().update(_names)

from pose_unsupervised_b200.core.inference import (  # noqa: E402,F401
    get_max_preds, get_final_preds, decode_heatmaps, decode_heatmaps_flip)
