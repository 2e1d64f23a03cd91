from pose_unsupervised_b200.multiviews.triangulate import (  # noqa: F401
    triangulate_poses, ransac, reproject_poses, lift_heatmaps)
