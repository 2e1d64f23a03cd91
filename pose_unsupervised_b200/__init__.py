"""pose_unsupervised_b200 -- the multiview 2D->3D lifting hot path of
LouisNUST/pose-unsupervised as hand-written sm_100a CUDA behind the reference's
Python call signatures.

    from pose_unsupervised_b200.core.inference import get_max_preds, get_final_preds
    from pose_unsupervised_b200.multiviews.cameras import project_pose, camera_to_world_frame
    from pose_unsupervised_b200.multiviews.triangulate import triangulate_poses, ransac, reproject_poses
    from pose_unsupervised_b200.multiviews.pictorial import rpsm
    from pose_unsupervised_b200.multiviews.body import HumanBody
    from pose_unsupervised_b200.core.loss import FundamentalLoss

or put ``pose_unsupervised_b200/dropin`` in front of the reference's ``lib`` on
``sys.path``.  All arithmetic runs in ``libposeb200.so``
(include/poseb200.h); there is no CPU fallback.
"""
from ._lib import Pb200Error, LIB_PATH  # noqa: F401

__version__ = '0.1.0'
