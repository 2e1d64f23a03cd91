from pose_unsupervised_b200.multiviews.body import HumanBody  # noqa: F401
