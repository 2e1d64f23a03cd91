#!/usr/bin/env python
"""bench.py -- multiview frames/s of the lifting hot path on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload lift|rpsm|pseudo] [--frames B] [--views V] [--joints J] [--hw S]

A "step" is one pass of the hot path over one batch of synthetic input.  The default
workload is BASELINE.json configs[1]: batched heatmap decode + DLT triangulation +
reprojection error, 4 views x 17 joints x 64x64 float32, 4096 frames per GPU (weak
scaling: every rank owns its own 4096 frames, the exchange step is an all-gather of
the 3D poses and an all-reduce of the MPJPE partial sums).

ONE JSON line is printed by rank 0:
  value        frames/s, whole job, inputs resident in HBM, CUDA-event time, max over ranks
  e2e          the same metric through the numpy-in / numpy-out public API
               (pose_unsupervised_b200.multiviews.triangulate.lift_heatmaps), host->device
               copy of the heatmaps from pinned memory and device->host copy of the
               results inside the timed region
  roofline     the dominant kernel (decode_tma_kernel, the only pass over the heatmaps) against the
               measured HBM peak; `lift_path_ms` / `whole_path_frac` cover decode + lift together
  cpu_baseline the oracle port of the reference's CPU path timed on this box's host cores
               on a bounded sample of the same workload (rank 0, N=1 only)
--impl reference times that CPU path alone (all host cores) and prints the same line shape.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

V, J, HW = 4, 17, 64                      # BASELINE.json configs[1]; --views/--joints/--hw run the sweep of configs[4]


def bytes_per_frame():
    """algorithmic bytes per frame of decode + triangulate + reproject
    (heatmaps read once + xy/maxval + center/scale + camera ids + X + reprojection error)."""
    return V * J * HW * HW * 4 + V * J * 12 + V * 16 + V * 8 + J * 24 + V * J * 4    # 1,115,704 at 4/17/64


def metric_name():
    return 'multiview frames/s (%d views x %d joints, %dx%d)' % (V, J, HW, HW)


def measured_hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------
def make_side_inputs(B, seed):
    """center/scale per row, 28-camera table (7 subjects x 4, as H36M), camera index per row."""
    from pose_unsupervised_b200.multiviews.cameras import pack_camera
    from pose_unsupervised_b200.utils import synth
    rng = np.random.default_rng(seed)
    rigs = synth.camera_table(7, V, seed=0)
    pack = np.array([pack_camera(c) for rig in rigs for c in rig])
    subj = rng.integers(0, 7, B)
    index = (subj[:, None] * V + np.arange(V)[None]).reshape(-1).astype(np.int32)
    center = rng.uniform(400, 600, (B * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * V, 1)), 2, axis=1)
    return rigs, subj, pack, index, center, scale


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
               0x4: 'sw_power_cap', 0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


# ---------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path (decode -> triangulate -> reproject)
# ---------------------------------------------------------------------------------------
def _cpu_frames(args):
    """Worker: the reference's CPU path on `n` frames (its own loops, its own cv2 call)."""
    seed, n = args
    import warnings
    warnings.filterwarnings('ignore')      # random heatmaps triangulate to far-away points
    from oracle import inference as oinf
    from oracle import transforms as otr
    from oracle import triangulate as otri
    from pose_unsupervised_b200.utils import synth
    try:
        import cv2  # noqa: F401
        otr.set_backend('cv2')             # what lib/utils/transforms.py:105-107 calls
    except Exception:
        otr.set_backend('lu')
    rng = np.random.default_rng(seed)
    rig = synth.camera_ring(V, seed=seed % 7)
    cams = [rig[v] for _ in range(n) for v in range(V)]
    hm = rng.random((n * V, J, HW, HW), dtype=np.float32)
    center = rng.uniform(400, 600, (n * V, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (n * V, 1)), 2, axis=1)
    t0 = time.perf_counter()
    preds, maxvals = oinf.get_final_preds_loops(True, hm, center, scale)        # lib/core/inference.py:50-75
    vis = np.ones(preds.shape[:2])
    proj, _ = otri.reproject_poses(preds, cams, vis, nviews=V)                             # lib/multiviews/triangulate.py:169-213
    _ = np.linalg.norm(proj - preds, axis=2)
    return time.perf_counter() - t0


def cpu_path_rate(frames_per_worker, workers, pool=None):
    """frames/s of the CPU path with `workers` processes each doing `frames_per_worker` frames."""
    jobs = [(1000 + w, frames_per_worker) for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        _cpu_frames(jobs[0])
    else:
        pool.map(_cpu_frames, jobs)
    return workers * frames_per_worker / (time.perf_counter() - t0)


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path, all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    import multiprocessing as mp
    cores = host_cores()
    steps = args.steps if args.steps else 3
    warmup = args.warmup if args.warmup is not None else 1
    steps, warmup = min(steps, 5), min(warmup, 2)        # bounded: each step is seconds of CPU work
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, warmup)):
            pilot = cpu_path_rate(4, cores, pool)           # also pays imports / page-in
        per_worker = int(min(2048, max(8, pilot / cores * 6)))   # ~6 s of CPU work per step
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_path_rate(per_worker, cores, pool)
        dt = time.perf_counter() - t0
    frames = steps * cores * per_worker
    value = frames / dt
    sample = '%d steps x %d frames (%d per process, %d processes) of the lift workload' % (
        steps, cores * per_worker, per_worker, cores)
    line = {
        'impl': 'reference', 'metric': metric_name(), 'value': value, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 decode, f64 lift', 'data': 'synthetic',
        'config': dict(workload_config(args.frames), sample_frames_per_step=cores * per_worker),
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(frames_per_gpu):
    return {'workload': 'configs[1]: batched heatmap decode + DLT triangulation + reprojection error',
            'views': V, 'joints': J, 'heatmap': '%dx%d float32' % (HW, HW), 'frames_per_gpu': frames_per_gpu,
            'cameras': '28-camera table (7 rigs x 4 views)', 'post_process': True,
            'l2': 'no flush: each step streams %.2f GB per GPU, far above the 126 MB L2'
                  % (frames_per_gpu * bytes_per_frame() / 1e9)}


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pose_unsupervised_b200 import parallel, runtime as rt
    from pose_unsupervised_b200.multiviews.cameras import CameraTable
    from pose_unsupervised_b200.core.inference import decode_heatmaps
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps, mpjpe_stats
    from pose_unsupervised_b200.utils.transforms import crop_affine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if args.lift_variant is not None:
        from pose_unsupervised_b200 import _lib
        _lib.call('pb200_set_tuning', 1, args.lift_variant)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    steps = args.steps if args.steps else 200
    warmup = args.warmup if args.warmup is not None else 10
    warmup = max(warmup, 3)
    B = args.frames

    rigs, subj, pack, index, center, scale = make_side_inputs(B, seed=rank)
    table = CameraTable.from_arrays(pack, index)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    hm = torch.rand((B * V, J, HW, HW), generator=g, device=dev, dtype=torch.float32)
    d_center, d_scale = rt.to_device(center), rt.to_device(scale)
    gt = torch.zeros((B, J, 3), dtype=torch.float64, device=dev)   # MPJPE reference for the exchange step
    nframes_total = B * world
    # exchange step: ONE all-gather of [3D poses | MPJPE partial sums]; the lift kernel writes the
    # poses straight into the send buffer
    exch = parallel.PoseExchange(nframes_total, J, dev)

    def step(ev=None):
        aff = crop_affine(d_center, d_scale, (HW, HW), inv=1)
        if ev is not None:
            ev[0].record()
        res = lift_heatmaps(hm, None, None, table, nviews=V, post_process=True, affine=aff,
                            out_poses3d=exch.poses_view())
        if ev is not None:
            ev[1].record()
        exch.stats_view().zero_()
        mpjpe_stats(res.poses3d, gt, out=exch.stats_view())
        if world > 1:
            exch.run()
        return res

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    fence()

    # The whole step (3 kernels, a memset and the NCCL all-gather) is captured once in a CUDA graph
    # and replayed: at ~0.8 ms per step the Python/launch overhead of the eager path is otherwise
    # visible, above all at N > 1.
    run_step, graphed = step, False
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            run_step, graphed = graph.replay, True
            for _ in range(3):
                run_step()
        except Exception as e:   # fall back to eager launches, say so in the JSON line
            sys.stderr.write('CUDA graph capture failed (%s); timing eager launches\n' % e)
            run_step, graphed = step, False
    fence()

    clocks = ClockSampler(local)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    fence()
    start.record()
    for i in range(steps):
        run_step()
    stop.record()
    fence()
    clock_info = clocks.stop()
    ms_total = parallel.max_over_ranks(start.elapsed_time(stop), dev)
    value = nframes_total * steps / (ms_total * 1e-3)
    variant = 2 if args.lift_variant is None else args.lift_variant
    # crop_affine_kernel, lift kernel(s), memset of the 4 sums, mpjpe_kernel
    launches_per_step = 5 if variant == 2 else 4

    # ---- the dominant kernel alone: CUDA events around every launch (eager pass, same work) ----
    ksteps = min(steps, 50)
    kernel_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(ksteps)]
    fence()
    for i in range(ksteps):
        step(kernel_events[i])
    fence()
    lift_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    lift_ms = parallel.max_over_ranks(lift_ms, dev)
    if variant == 2:
        # the lift is two kernels; the dominant one (decode_tma_kernel, the only pass over the
        # heatmaps) is timed alone through the decode entry point, same inputs, same launch
        aff = crop_affine(d_center, d_scale, (HW, HW), inv=1)
        fence()
        for i in range(ksteps):
            kernel_events[i][0].record()
            decode_heatmaps(hm, post_process=True, affine=aff)
            kernel_events[i][1].record()
        fence()
        kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
        kern_ms = parallel.max_over_ranks(kern_ms, dev)
        kern_name = 'decode_tma_kernel'
        kern_bytes = B * (V * J * HW * HW * 4 + V * J * 12 + V * 48)   # heatmaps + xy/maxval + affine rows
    else:
        kern_ms = lift_ms
        kern_name = 'lift_fused_kernel' if variant == 0 else 'lift_fused_tma_kernel'
        kern_bytes = B * bytes_per_frame()

    # ---- end to end through the public numpy API, host buffers ------------------------------
    e2e_steps = max(1, min(3, steps))
    pinned = torch.empty((B * V, J, HW, HW), dtype=torch.float32, pin_memory=True)
    pinned.copy_(hm)
    hm_host = pinned.numpy()                           # numpy view of pinned memory
    cams_arg = table                                   # built once, like the reference's camera list
    torch.cuda.synchronize()
    d2h = 0

    def e2e_step():
        res = lift_heatmaps(hm_host, center, scale, cams_arg, nviews=V, post_process=True).numpy()
        return res

    e2e_step()                                         # warm-up (allocator, page mapping)
    fence()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = parallel.max_over_ranks(e2e_s, dev)
    d2h = sum(a.nbytes for a in (out.xy, out.maxvals, out.poses3d, out.reproj_err))
    h2d = hm_host.nbytes + center.nbytes + scale.nbytes
    e2e_value = nframes_total * e2e_steps / e2e_s

    # ---- sanity: the timed path is the real path (spot check against the oracle on rank 0) -----
    res = step()
    torch.cuda.synchronize()

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_leg()
        line = {
            'metric': metric_name(), 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms_total / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 decode, f64 lift', 'data': 'synthetic',
            'config': dict(workload_config(B), cuda_graph=graphed),
            'clocks': clock_info,
            'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps,
                    'api': 'pose_unsupervised_b200.multiviews.triangulate.lift_heatmaps (numpy in, numpy out)'},
            'gpu_launches': steps * launches_per_step,
            'roofline': {'bound': 'hbm', 'kernel': kern_name, 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                         'peak_source': peak_src, 'kernel_ms': kern_ms,
                         'algorithmic_bytes_per_launch': kern_bytes,
                         'lift_path_ms': lift_ms, 'lift_variant': variant,
                         'whole_path_frac': B * bytes_per_frame() / (lift_ms * 1e-3) / 1e9 / peak},
            'cpu_baseline': cpu,
        }
        traffic = os.path.join(ROOT, 'profiles', 'lift_fused_traffic.json')
        if os.path.exists(traffic):
            try:
                with open(traffic) as f:
                    t = json.load(f)
                if t.get('frames_per_launch') == B and t.get('kernel') == kern_name:
                    line['roofline']['traffic'] = t.get('dram_bytes_per_launch')
            except Exception:
                pass
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_leg():
    """~10-30 s of the oracle port on all host cores (and the single-core figure)."""
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    import multiprocessing as mp
    cores = host_cores()
    cpu_path_rate(1, 1)                                        # imports, page-in
    one = cpu_path_rate(16, 1)                                 # pilot, also the per-core number
    per_worker = int(min(4096, max(8, one * 12)))              # ~12 s per process
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        rate = cpu_path_rate(per_worker, cores, pool)
    return {'value': rate, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'single_core_value': one,
            'sample': '%d frames (%d per process x %d processes) of the same workload: reference loops '
                      'of get_final_preds + reproject_poses (oracle port; pymvg restated)'
                      % (per_worker * cores, per_worker, cores)}


# ---------------------------------------------------------------------------------------
# secondary workloads (not the headline line; used for profiles/ and DESIGN.md numbers)
# ---------------------------------------------------------------------------------------
def run_rpsm(args):
    import torch
    from pose_unsupervised_b200.multiviews import pictorial
    from pose_unsupervised_b200.multiviews.body import HumanBody
    from pose_unsupervised_b200.utils import synth
    import types
    torch.cuda.set_device(0)
    B = args.frames if args.frames != 4096 else 296
    body = HumanBody.h36m17()
    edges = body.edges()
    cfg = types.SimpleNamespace(
        NETWORK=types.SimpleNamespace(IMAGE_SIZE=np.array([256, 256]), HEATMAP_SIZE=np.array([64, 64])),
        PICT_STRUCT=types.SimpleNamespace(FIRST_NBINS=16, RECUR_NBINS=2, RECUR_DEPTH=10, GRID_SIZE=2000,
                                          LIMB_LENGTH_TOLERANCE=150))
    base = 8
    poses = synth.random_poses(base, seed=1)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)]))
           for e in edges}
    table = pictorial.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    hms, cams, centers, scales, roots, limbs = [], [], [], [], [], []
    for f in range(base):
        cam = synth.camera_ring(4, seed=50 + f)
        boxes = synth.crop_box(cam, poses[f])
        hms.append(synth.gaussian_heatmaps(cam, boxes, poses[f], 64, 256, 2.0, 0.02, seed=f))
        cams.append(cam)
        roots.append(poses[f][0] + [20.0, -30.0, 10.0])
        centers.append([b['center'] for b in boxes])
        scales.append([b['scale'] for b in boxes])
        limb = synth.limb_lengths(poses[f], edges)
        limbs.append([limb[e] for e in edges])
    rep = [i % base for i in range(B)]
    hm = torch.from_numpy(np.array(hms)).cuda()[rep].contiguous()
    cam_list = [c for i in rep for c in cams[i]]
    from pose_unsupervised_b200.multiviews.cameras import CameraTable
    ctab = CameraTable.from_cameras(cam_list)
    cen = np.array([centers[i] for i in rep]).reshape(-1, 2)
    sca = np.array([scales[i] for i in rep]).reshape(-1, 2)
    roo = np.array([roots[i] for i in rep])
    lim = np.array([limbs[i] for i in rep])
    steps = args.steps if args.steps else 3
    for _ in range(2):
        out = pictorial.rpsm_batch(ctab, hm, cen, sca, roo, lim, table, cfg, body)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        out = pictorial.rpsm_batch(ctab, hm, cen, sca, roo, lim, table, cfg, body)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    err = float(np.mean(np.linalg.norm(out.cpu().numpy()[:base] - poses, axis=2)))
    print(json.dumps({'workload': 'configs[2]: RPSM 4 views x 17 joints, 16^3 then 10 x 2^3', 'frames': B,
                      'ms_per_step': ms, 'frames_per_s': B / (ms * 1e-3), 'mpjpe_mm_vs_synthetic_gt': err}))


def run_pseudo(args):
    """BASELINE.json configs[3] variant (ii): the pseudo-label pass of run/test/test_pseudo_label.py
    from 2D locations -- confidence threshold, RANSAC view selection, triangulate + reproject,
    epipolar residuals -- on `--frames` frames in total, sharded by frame over the ranks."""
    import torch
    import torch.distributed as dist
    from pose_unsupervised_b200 import parallel, runtime as rt
    from pose_unsupervised_b200.core.loss import FundamentalTable, epipolar_residuals
    from pose_unsupervised_b200.multiviews.cameras import CameraTable, pack_camera
    from pose_unsupervised_b200.multiviews.triangulate import ransac, reproject_poses
    from pose_unsupervised_b200.utils import synth
    import types
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    total = args.frames if args.frames != 4096 else 1000000
    lo, hi = parallel.frame_shard(total, rank, world)
    B = hi - lo
    rng = np.random.default_rng(100 + rank)
    rigs = synth.camera_table(7, 4, seed=0)
    pack = np.array([pack_camera(c) for rig in rigs for c in rig])
    subj = rng.integers(0, 7, B)
    base = synth.random_poses(1024, seed=5)
    poses = base[rng.integers(0, 1024, B)] + rng.normal(0, 15, (B, 17, 3))
    obs = np.empty((B * 4, 17, 2))
    for s_ in range(7):
        sel = np.where(subj == s_)[0]
        for v in range(4):
            obs[sel * 4 + v] = synth.project_plumb_bob_numpy(poses[sel].reshape(-1, 3), rigs[s_][v]).reshape(len(sel), 17, 2)
    obs += rng.normal(0, 2.0, obs.shape)
    bad = rng.random(obs.shape[:2]) < 0.10
    obs[bad] += rng.normal(0, 50.0, (int(bad.sum()), 2))
    conf = rng.uniform(0.04, 1.12, obs.shape[:2]).astype(np.float32)
    table = CameraTable.from_arrays(pack, (subj[:, None] * 4 + np.arange(4)[None]).reshape(-1))
    ftab = FundamentalTable.from_cameras({s_: rigs[s_] for s_ in range(7)})
    d_obs = rt.to_device(obs.astype(np.float32))
    d_conf = rt.to_device(conf)
    d_subj = torch.from_numpy(subj).to(dev)
    cfg = types.SimpleNamespace(DATASET=types.SimpleNamespace(NO_DISTORTION=False),
                                PSEUDO_LABEL=types.SimpleNamespace(REPROJ_THRE=10.0, NUM_INLIERS=3))
    subj_list = ftab.slots(subj)                                      # table slots, computed once

    def step():
        vis = d_conf > 0.7                                            # test_pseudo_label.py:194
        vis = ransac(d_obs, table, vis, cfg)                          # :221
        proj, pvis, pts = reproject_poses(d_obs, table, vis, False, return_points=True)   # :237
        resid = epipolar_residuals(proj, subj_list, ftab)             # test_fund_mtx.py:56-69 on the labels
        if world > 1:
            parallel.gather_poses(pts, total)
        return proj, pvis, pts, resid

    for _ in range(2):
        out = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = args.steps if args.steps else 5
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        out = step()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(a.elapsed_time(b) / steps, dev)
    if rank == 0:
        proj, pvis, pts, resid = out
        keep = float(pvis.float().mean())
        err = float((pts.cpu().numpy() - poses)[pvis.view(B, 4, 17)[:, 0].cpu().numpy() > 0].__abs__().mean()) if B else 0.0
        per_frame = 4 * 17 * 12 + 32 + 4 * 17 * 8 + 4 * 17 + 17 * 24 + 12 * 17 * 8
        print(json.dumps({'workload': 'configs[3](ii): pseudo-label pass from 2D locations (conf>0.7, RANSAC 3 inliers/10 px, '
                                      'reproject, epipolar residuals)', 'frames': total, 'n_gpus': world, 'ms_per_step': ms,
                          'frames_per_s': total / (ms * 1e-3), 'algorithmic_GBps': total * per_frame / (ms * 1e-3) / 1e9,
                          'labels_kept': keep, 'mean_abs_3d_err_mm': err}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=0)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='lift', choices=['lift', 'rpsm', 'pseudo'])
    ap.add_argument('--views', type=int, default=4)
    ap.add_argument('--joints', type=int, default=17)
    ap.add_argument('--hw', type=int, default=64, help='heatmap side')
    ap.add_argument('--frames', type=int, default=4096, help='frames per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph replay')
    ap.add_argument('--lift-variant', type=int, default=None, help='0 = LDG front end, 1 = TMA ring (default)')
    args = ap.parse_args()
    global V, J, HW
    V, J, HW = args.views, args.joints, args.hw
    if args.impl == 'reference':
        run_reference_arm(args)
    elif args.workload == 'rpsm':
        run_rpsm(args)
    elif args.workload == 'pseudo':
        run_pseudo(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
