from pose_unsupervised_b200.core.inference import *  # noqa: F401,F403
from pose_unsupervised_b200.core.inference import get_max_preds, get_final_preds, decode_heatmaps  # noqa: F401
