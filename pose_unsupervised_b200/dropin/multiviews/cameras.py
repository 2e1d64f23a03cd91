from pose_unsupervised_b200.multiviews.cameras import (  # noqa: F401
    unfold_camera_param, project_pose, world_to_camera_frame, camera_to_world_frame, CameraTable)
