"""H36M camera model behind the names of lib/multiviews/cameras.py:12-82.

Camera dicts (``R T fx fy cx cy k p``) are packed into a float64 table on the
device (``CameraTable``); every function that takes the reference's python list of
per-row camera dicts also accepts a ``CameraTable`` built once by the caller.
"""
import numpy as np
import torch

from .. import _lib, runtime as rt


def pack_camera(camera):
    """dict -> the PB200_CAM_STRIDE float64 record of include/poseb200.h."""
    rec = np.zeros(_lib.CAM_STRIDE)
    rec[0:9] = np.asarray(camera['R'], dtype=np.float64).reshape(9)
    rec[9:12] = np.asarray(camera['T'], dtype=np.float64).reshape(3)
    rec[12] = np.asarray(camera['fx'], dtype=np.float64).reshape(-1)[0]
    rec[13] = np.asarray(camera['fy'], dtype=np.float64).reshape(-1)[0]
    rec[14] = np.asarray(camera['cx'], dtype=np.float64).reshape(-1)[0]
    rec[15] = np.asarray(camera['cy'], dtype=np.float64).reshape(-1)[0]
    rec[16:19] = np.asarray(camera['k'], dtype=np.float64).reshape(3)
    rec[19:21] = np.asarray(camera['p'], dtype=np.float64).reshape(2)
    return rec


class CameraTable(object):
    """Packed cameras on the device + the camera id of every (frame, view) row."""

    def __init__(self, pack, index):
        self.pack = pack          # CUDA float64 [ncam, 24]
        self.index = index        # CUDA int32 [nrows]

    def __len__(self):
        return int(self.index.shape[0])

    @classmethod
    def from_cameras(cls, camera_params):
        """List of per-row camera dicts (lib/multiviews/triangulate.py:62-63) -> table.

        Identical dict objects are stored once (the H36M db repeats 28 calibrations).
        """
        if isinstance(camera_params, CameraTable):
            return camera_params
        seen, recs = {}, []
        index = np.empty(len(camera_params), dtype=np.int32)
        for i, cam in enumerate(camera_params):
            k = id(cam)
            slot = seen.get(k)
            if slot is None:
                slot = len(recs)
                seen[k] = slot
                recs.append(pack_camera(cam))
            index[i] = slot
        pack = np.stack(recs) if recs else np.zeros((0, _lib.CAM_STRIDE))
        return cls(rt.to_device(pack), rt.to_device(index))

    @classmethod
    def from_arrays(cls, pack, index):
        """pack [ncam,24] float64, index [nrows] int -> table (arrays or CUDA tensors)."""
        return cls(rt.to_device(pack, torch.float64).reshape(-1, _lib.CAM_STRIDE),
                   rt.to_device(index, torch.int32).reshape(-1))


def unfold_camera_param(camera, avg_f=True):
    """lib/multiviews/cameras.py:12-22 (host-side dict access, no arithmetic on the path)."""
    if avg_f:
        f = 0.5 * (camera['fx'] + camera['fy'])
    else:
        f = np.array([camera['fx'], camera['fy']])
    c = np.array([camera['cx'], camera['cy']])
    return camera['R'], camera['T'], f, c, camera['k'], camera['p']


def _project(x, camera, model):
    rt.require_device()
    pack = rt.to_device(pack_camera(camera)[None])
    pts = rt.to_device(x, torch.float64).reshape(-1, 3)
    out = rt.empty((pts.shape[0], 2), torch.float64)
    _lib.call('pb200_project', rt.ptr(pack), 0, rt.ptr(pts), pts.shape[0], model, rt.ptr(out),
              rt.stream_ptr())
    return out if rt.is_device_tensor(x) else rt.to_host(out)


def project_point_radial(x, R, T, f, c, k, p):
    """lib/multiviews/cameras.py:25-49 on unfolded parameters: ``f`` is the averaged focal length
    (shape (1,), what ``project_pose`` passes) or ``[fx, fy]`` (``unfold_camera_param(avg_f=False)``)."""
    f = np.asarray(f, dtype=np.float64).reshape(-1)
    c = np.asarray(c, dtype=np.float64).reshape(-1)
    camera = {'R': R, 'T': T, 'fx': f[:1], 'fy': f[-1:], 'cx': c[:1], 'cy': c[1:2], 'k': k, 'p': p}
    return _project(x, camera, 0 if f.size == 1 else 3)


def project_pose(x, camera):
    """lib/multiviews/cameras.py:25-54: world [n,3] -> pixels [n,2], averaged focal length."""
    return _project(x, camera, 0)


def project_pose_plumb_bob(x, camera, distorted=True):
    """pymvg find2d model (lib/multiviews/triangulate.py:147,210): separate fx/fy, OpenCV distortion."""
    return _project(x, camera, 1 if distorted else 2)


def _frame_change(x, R, T, to_world):
    rt.require_device()
    pts = rt.to_device(x, torch.float64).reshape(-1, 3)
    r = rt.to_device(np.asarray(R, dtype=np.float64).reshape(9))
    t = rt.to_device(np.asarray(T, dtype=np.float64).reshape(3))
    out = rt.empty(pts.shape, torch.float64)
    _lib.call('pb200_frame_change', rt.ptr(r), rt.ptr(t), rt.ptr(pts), pts.shape[0], to_world,
              rt.ptr(out), rt.stream_ptr())
    return out if rt.is_device_tensor(x) else rt.to_host(out)


def world_to_camera_frame(x, R, T):
    """lib/multiviews/cameras.py:57-68:  R (x - T)."""
    return _frame_change(x, R, T, 0)


def camera_to_world_frame(x, R, T):
    """lib/multiviews/cameras.py:71-82:  R^T x + T."""
    return _frame_change(x, R, T, 1)
