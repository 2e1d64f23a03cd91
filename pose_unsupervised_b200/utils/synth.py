"""Synthetic inputs for the lifting path: H36M-style camera rigs, poses, heatmaps.

No dataset ships with the reference, so tests, the
golden-vector script and bench.py all draw their inputs from here.  Everything
is seeded numpy; nothing here is part of the compute path (the small numpy
pin-hole projection below only *renders inputs*).

Camera dict layout = what lib/multiviews/cameras.py:12-22 and
lib/multiviews/triangulate.py:28 read: ``R [3,3]``, ``T [3,1]`` (camera centre,
world mm), ``fx fy cx cy`` shape-(1,), ``k [3,1]``, ``p [2,1]``.
"""
import numpy as np

# canonical standing pose, H36M 17-joint order
# (lib/dataset/multiview_h36m_compatible.py:26-44), world mm, z up
H36M17_REST = np.array([
    [0, 0, 920],        # root
    [-130, 0, 920],     # rhip
    [-135, 10, 480],    # rkne
    [-135, 0, 60],      # rank
    [130, 0, 920],      # lhip
    [135, 10, 480],     # lkne
    [135, 0, 60],       # lank
    [0, -10, 1160],     # belly
    [0, 0, 1420],       # neck
    [0, 70, 1520],      # nose
    [0, 10, 1630],      # head
    [170, 0, 1400],     # lsho
    [230, 20, 1120],    # lelb
    [250, 120, 880],    # lwri
    [-170, 0, 1400],    # rsho
    [-230, 20, 1120],   # relb
    [-250, 120, 880],   # rwri
], dtype=np.float64)

# H36M-17 index of each joint of the reference's 16-joint tree
# (lib/multiviews/body.py:22-26: rank rkne rhip lhip lkne lank root thorax
#  upper-neck head-top rwri relb rsho lsho lelb lwri)
MPII16_FROM_H36M17 = [3, 2, 1, 4, 5, 6, 0, 8, 9, 10, 16, 15, 14, 11, 12, 13]


def look_at_camera(centre, target, rng=None, jitter=True):
    """One H36M-style camera at ``centre`` looking at ``target`` (world z up, image y down)."""
    centre = np.asarray(centre, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - centre
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])
    z = np.zeros(1)
    if rng is None or not jitter:
        fx, fy, cx, cy = 1145.0, 1144.0, 512.0, 515.0
        k = np.array([-0.2, 0.24, -0.002])
        p = np.array([-9e-4, 6e-4])
    else:
        fx = 1145.0 + rng.normal(0, 3)
        fy = 1145.0 + rng.normal(0, 3)
        cx = 512.0 + rng.normal(0, 6)
        cy = 515.0 + rng.normal(0, 6)
        k = np.array([-0.2, 0.24, -0.002]) * (1 + rng.normal(0, 0.05, 3))
        p = np.array([-9e-4, 6e-4]) * (1 + rng.normal(0, 0.05, 2))
    return {'R': R, 'T': centre.reshape(3, 1),
            'fx': z + fx, 'fy': z + fy, 'cx': z + cx, 'cy': z + cy,
            'k': k.reshape(3, 1), 'p': p.reshape(2, 1)}


def camera_ring(nviews=4, seed=0, radius=4500.0, height=1500.0, target=(0.0, 0.0, 900.0)):
    """``nviews`` cameras on a ring looking at ``target``."""
    rng = np.random.default_rng(seed)
    cams = []
    phase = rng.uniform(0, 2 * np.pi)
    for v in range(nviews):
        ang = phase + 2 * np.pi * v / nviews + rng.normal(0, 0.08)
        r = radius * (1 + rng.normal(0, 0.05))
        centre = [r * np.cos(ang), r * np.sin(ang), height * (1 + rng.normal(0, 0.05))]
        cams.append(look_at_camera(centre, target, rng))
    return cams


def camera_table(nsubjects=7, nviews=4, seed=0):
    """``nsubjects`` rigs (H36M has 7 subjects x 4 cameras = 28 calibrations)."""
    return [camera_ring(nviews, seed=seed * 1000 + s) for s in range(nsubjects)]


def random_poses(nframes, seed=0, njoints=17, spread=400.0, joint_sigma=60.0):
    """[nframes, njoints, 3] world-mm poses: rest pose, random yaw, offset and joint noise."""
    rng = np.random.default_rng(seed)
    out = np.empty((nframes, 17, 3))
    for i in range(nframes):
        yaw = rng.uniform(0, 2 * np.pi)
        c, s = np.cos(yaw), np.sin(yaw)
        rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
        p = H36M17_REST @ rot.T + rng.normal(0, joint_sigma, (17, 3))
        p[:, :2] += rng.uniform(-spread, spread, 2)
        out[i] = p
    if njoints == 17:
        return out
    if njoints == 16:
        return out[:, MPII16_FROM_H36M17]
    idx = np.arange(njoints) % 17
    return out[:, idx] + rng.normal(0, 5.0, (nframes, njoints, 3))


def project_h36m_numpy(pts, cam):
    """Input renderer: averaged-f H36M projection (same model as cameras.project_pose)."""
    k = np.reshape(cam['k'], 3)
    p = np.reshape(cam['p'], 2)
    xc = cam['R'] @ (pts.T - cam['T'])
    u, v = xc[0] / xc[2], xc[1] / xc[2]
    r2 = u * u + v * v
    gain = 1 + k[0] * r2 + k[1] * r2 ** 2 + k[2] * r2 ** 3 + p[0] * v + p[1] * u
    f = 0.5 * (cam['fx'][0] + cam['fy'][0])
    return np.stack([f * (u * gain + p[1] * r2) + cam['cx'][0],
                     f * (v * gain + p[0] * r2) + cam['cy'][0]], axis=1)


def project_plumb_bob_numpy(pts, cam, distorted=True):
    """Input renderer: OpenCV plumb-bob projection with separate fx, fy."""
    k = np.reshape(cam['k'], 3)
    p = np.reshape(cam['p'], 2)
    xc = cam['R'] @ (pts.T - cam['T'])
    x, y = xc[0] / xc[2], xc[1] / xc[2]
    if distorted:
        r2 = x * x + y * y
        barrel = 1 + k[0] * r2 + k[1] * r2 ** 2 + k[2] * r2 ** 3
        xd = x * barrel + 2 * p[0] * x * y + p[1] * (r2 + 2 * x * x)
        yd = y * barrel + p[0] * (r2 + 2 * y * y) + 2 * p[1] * x * y
        x, y = xd, yd
    return np.stack([cam['fx'][0] * x + cam['cx'][0], cam['fy'][0] * y + cam['cy'][0]], axis=1)


def multiview_observations(poses, rigs, rig_of_frame, noise_px=0.0, outlier_frac=0.0,
                           outlier_px=50.0, seed=0, distorted=True):
    """2D observations [B*V, J, 2] (view-minor rows) + the per-row camera list."""
    rng = np.random.default_rng(seed)
    nframes, njoints = poses.shape[:2]
    nviews = len(rigs[0])
    obs = np.empty((nframes * nviews, njoints, 2))
    cams = []
    for i in range(nframes):
        rig = rigs[rig_of_frame[i]]
        for v in range(nviews):
            obs[i * nviews + v] = project_plumb_bob_numpy(poses[i], rig[v], distorted)
            cams.append(rig[v])
    if noise_px > 0:
        obs += rng.normal(0, noise_px, obs.shape)
    if outlier_frac > 0:
        bad = rng.random(obs.shape[:2]) < outlier_frac
        obs[bad] += rng.normal(0, outlier_px, (int(bad.sum()), 2))
    return obs, cams


def crop_box(cams, pose, pad=1.25):
    """Per-view crop {center [2], scale [2]} (scale*200 px square) around the projected pose."""
    boxes = []
    for cam in cams:
        xy = project_h36m_numpy(pose, cam)
        lo, hi = xy.min(0), xy.max(0)
        side = float(max(hi - lo)) * pad
        s = side / 200.0
        boxes.append({'center': (0.5 * (lo + hi)).astype(np.float64),
                      'scale': np.array([s, s], dtype=np.float64)})
    return boxes


def crop_affine_numpy(center, scale0, out_w, out_h):
    """Input renderer: rot=0 crop map image px -> crop px (plain float64, no f32 staging)."""
    s = out_w / (scale0 * 200.0)
    return np.array([[s, 0, out_w * 0.5 - s * center[0]],
                     [0, s, out_h * 0.5 - s * center[1]]])


def gaussian_heatmaps(cams, boxes, pose, hm_size=64, img_size=256, sigma=2.0,
                      noise=0.02, seed=0):
    """[V, J, hm, hm] float32: Gaussians at the projected joints + a uniform noise floor."""
    rng = np.random.default_rng(seed)
    nviews, njoints = len(cams), pose.shape[0]
    ys, xs = np.mgrid[0:hm_size, 0:hm_size].astype(np.float64)
    out = np.empty((nviews, njoints, hm_size, hm_size), dtype=np.float32)
    for v, (cam, box) in enumerate(zip(cams, boxes)):
        t = crop_affine_numpy(box['center'], box['scale'][0], img_size, img_size)
        xy = project_h36m_numpy(pose, cam)
        xy = (xy @ t[:, :2].T + t[:, 2]) * hm_size / img_size
        for j in range(njoints):
            g = np.exp(-((xs - xy[j, 0]) ** 2 + (ys - xy[j, 1]) ** 2) / (2 * sigma ** 2))
            out[v, j] = (g + noise * rng.random((hm_size, hm_size))).astype(np.float32)
    return out


def limb_lengths(pose, edges):
    """{(parent, child): mm} from a pose (run/test/test_rpsm.py:34-45)."""
    return {(p, c): float(np.linalg.norm(pose[p] - pose[c])) for p, c in edges}
