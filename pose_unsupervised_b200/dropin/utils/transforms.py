from pose_unsupervised_b200.utils.transforms import (  # noqa: F401
    get_affine_transform, affine_transform, transform_preds, crop_affine)
