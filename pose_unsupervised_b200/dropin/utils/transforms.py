"""utils.transforms with the overlay in front: everything the reference defines
(lib/utils/transforms.py: flip_back, fliplr_joints, crop, flip_back_th, ...) stays reachable;
the functions on the lifting path are replaced by this repository's.

``get_affine_transform`` and ``affine_transform`` stay the reference's own when its file is
present: the dataset calls them once per sample inside forked DataLoader workers
(lib/dataset/joints_dataset_compatible.py:161,177), where a CUDA round trip is neither possible
nor useful.  The lifting path itself never calls them -- it uses the batched device form
``crop_affine`` (the same arithmetic, bit for bit, rotations included).
"""
import utils as _pkg
from pose_unsupervised_b200.dropin._fallthrough import reference_names as _reference_names

_names, _reference = _reference_names(_pkg, 'transforms', __file__)
globals().update(_names)

from pose_unsupervised_b200.utils.transforms import (  # noqa: E402,F401
    crop_affine, transform_preds, generate_integral_preds_2d_th, transform_back_th)

if _reference is None:           # stand-alone: no reference file to fall through to
    from pose_unsupervised_b200.utils.transforms import get_affine_transform, affine_transform  # noqa: F401
