from pose_unsupervised_b200.multiviews.pictorial import rpsm, rpsm_batch, PairwiseTable  # noqa: F401
