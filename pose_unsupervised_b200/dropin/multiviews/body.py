"""multiviews.body with the overlay in front (lib/multiviews/body.py:11-57)."""
from pose_unsupervised_b200.multiviews.body import HumanBody  # noqa: F401
