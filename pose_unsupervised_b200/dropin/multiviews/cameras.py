"""multiviews.cameras with the overlay in front (lib/multiviews/cameras.py:12-82)."""
import multiviews as _pkg
from pose_unsupervised_b200.dropin._fallthrough import reference_names as _reference_names

_names, _reference = _reference_names(_pkg, 'cameras', __file__)
globals().update(_names)

from pose_unsupervised_b200.multiviews.cameras import (  # noqa: E402,F401
    unfold_camera_param, project_point_radial, project_pose, project_pose_plumb_bob,
    world_to_camera_frame, camera_to_world_frame, CameraTable, pack_camera)
