#!/usr/bin/env python
"""bench.py -- multiview frames/s of the lifting hot path on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload lift|rpsm|pseudo|pseudo-hm|sweep] [--frames B] [--views V] [--joints J] [--hw S]

A "step" is one pass of the hot path over one batch of synthetic input.  The default
workload is BASELINE.json configs[1]: batched heatmap decode + DLT triangulation +
reprojection error, 4 views x 17 joints x 64x64 float32, 4096 frames per GPU (weak
scaling: every rank owns its own 4096 frames; the exchange step is ONE all-gather of
[3D poses | MPJPE partial sums], double-buffered so that the all-gather of step k-1
runs on a side stream underneath the kernels of step k).

ONE JSON line is printed by rank 0:
  value        frames/s, whole job, inputs resident in HBM, CUDA-event time, max over ranks
  e2e          the same metric through the numpy-in / numpy-out public API
               (pose_unsupervised_b200.multiviews.triangulate.lift_heatmaps), host->device
               copy of the heatmaps from PINNED memory and device->host copy of the
               results inside the timed region; e2e_pageable is the same call on a plain
               (pageable) numpy array, which the API stages through pinned buffers in chunks
  roofline     the dominant kernel (decode_tma_kernel, the only pass over the heatmaps) against the
               measured HBM peak; `lift_path_ms` / `whole_path_frac` cover decode + lift together
  cpu_baseline the oracle port of the reference's CPU path timed on this box's host cores
               on a bounded sample of the same workload (rank 0, N=1 only)
  verified     the timed buffers were checked: argmax/maxval against torch on a slice, 3D
               poses against the oracle on 16 frames, and (N>1) the gathered poses against
               every rank's own shard
  secondary    (N=1) configs[2] RPSM with its own CPU baseline, configs[3](ii) pseudo-label pass
               over 1 M frames, and the decode+lift sweep of configs[4] on one GPU
--impl reference times the CPU path alone (all host cores) and prints the same line shape.
"""
import argparse
import gc
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


def bytes_per_frame(v=None, j=None, hw=None):
    """algorithmic bytes per frame of decode + triangulate + reproject
    (heatmaps read once + xy/maxval + center/scale + camera ids + X + reprojection error)."""
    v, j, hw = v or V, j or J, hw or HW
    return v * j * hw * hw * 4 + v * j * 12 + v * 16 + v * 8 + j * 24 + v * j * 4    # 1,115,704 at 4/17/64


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
def make_side_inputs(B, seed, nviews=None):
    """center/scale per row, 28-camera table (7 subjects x 4, as H36M), camera index per row."""
    from pose_unsupervised_b200.multiviews.cameras import pack_camera
    from pose_unsupervised_b200.utils import synth
    nv = nviews or V
    rng = np.random.default_rng(seed)
    rigs = synth.camera_table(7, nv, seed=0)
    pack = np.array([pack_camera(c) for rig in rigs for c in rig])
    subj = rng.integers(0, 7, B)
    index = (subj[:, None] * nv + np.arange(nv)[None]).reshape(-1).astype(np.int32)
    center = rng.uniform(400, 600, (B * nv, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (B * nv, 1)), 2, axis=1)
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
    seed, n, v, j, hw = args
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
    rig = synth.camera_ring(v, seed=seed % 7)
    cams = [rig[k] for _ in range(n) for k in range(v)]
    hm = rng.random((n * v, j, hw, hw), dtype=np.float32)
    center = rng.uniform(400, 600, (n * v, 2))
    scale = np.repeat(rng.uniform(1.5, 3.0, (n * v, 1)), 2, axis=1)
    t0 = time.perf_counter()
    preds, maxvals = oinf.get_final_preds_loops(True, hm, center, scale)        # lib/core/inference.py:50-75
    vis = np.ones(preds.shape[:2])
    proj, _ = otri.reproject_poses(preds, cams, vis, nviews=v)                  # lib/multiviews/triangulate.py:169-213
    _ = np.linalg.norm(proj - preds, axis=2)
    return time.perf_counter() - t0


def cpu_path_rate(frames_per_worker, workers, pool=None, shape=None):
    """frames/s of the CPU path with `workers` processes each doing `frames_per_worker` frames."""
    v, j, hw = shape or (V, J, HW)
    jobs = [(1000 + w, frames_per_worker, v, j, hw) for w in range(workers)]
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
    """--impl reference: the reference's CPU implementation of the path, all host cores.
    --steps / --warmup are honoured; each step is a bounded sample sized so that the whole run
    stays under REF_BUDGET_S seconds of wall clock."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    import multiprocessing as mp
    REF_BUDGET_S = 150.0
    cores = host_cores()
    steps = args.steps if args.steps else 3
    warmup = args.warmup if args.warmup is not None else 1
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pilot = cpu_path_rate(4, cores, pool)               # also pays imports / page-in
        per_step_s = max(0.25, min(6.0, REF_BUDGET_S / (steps + warmup)))
        per_worker = int(min(2048, max(2, pilot / cores * per_step_s)))
        for _ in range(warmup):
            cpu_path_rate(per_worker, cores, pool)
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


def cpu_baseline_leg(shape=None, seconds=12.0):
    """~10-30 s of the oracle port on all host cores (and the single-core figure)."""
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    import multiprocessing as mp
    cores = host_cores()
    cpu_path_rate(1, 1, shape=shape)                                        # imports, page-in
    one = cpu_path_rate(16, 1, shape=shape)                                 # pilot, also the per-core number
    per_worker = int(min(4096, max(4, one * seconds)))
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        rate = cpu_path_rate(per_worker, cores, pool, shape=shape)
    return {'value': rate, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'single_core_value': one,
            'sample': '%d frames (%d per process x %d processes) of the same workload: reference loops '
                      'of get_final_preds + reproject_poses (oracle port; pymvg restated)'
                      % (per_worker * cores, per_worker, cores)}


# ---------------------------------------------------------------------------------------
# our arm: configs[1]
# ---------------------------------------------------------------------------------------
def time_lift_device(hm, d_center, d_scale, table, B, nviews, steps, warmup=3):
    """ms per (crop affine + decode + lift) on resident inputs, CUDA events, eager launches."""
    import torch
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    from pose_unsupervised_b200.utils.transforms import crop_affine
    hw = int(hm.shape[-1])

    def one():
        aff = crop_affine(d_center, d_scale, (hw, hw), inv=1)
        return lift_heatmaps(hm, None, None, table, nviews=nviews, post_process=True, affine=aff)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        one()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def verify_lift(hm, pack, index, center, scale, aff, own_poses, exch, last_slot, rank, world):
    """The timed buffers hold the real path's results: argmax / maxval of a slice against torch,
    poses of 16 frames against the oracle (the checker), gathered poses against the local shard."""
    import torch
    import warnings
    from pose_unsupervised_b200.multiviews.cameras import CameraTable
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    frames = 16
    rows = frames * V
    sub = CameraTable(torch.from_numpy(pack).to(hm.device), torch.from_numpy(index[:rows]).to(hm.device))
    chk = lift_heatmaps(hm[:rows], None, None, sub, nviews=V, post_process=True, affine=aff[:rows],
                        return_idx=True)
    flat = hm[:rows].reshape(rows, J, -1)
    ok_idx = bool(torch.equal(chk.idx.long(), flat.argmax(dim=2)))
    ok_max = bool(torch.equal(chk.maxvals, flat.amax(dim=2)))
    ok_same = bool(torch.equal(chk.poses3d, own_poses[:frames]))                 # the graph wrote the same bits
    info = {'argmax_vs_torch': ok_idx, 'maxval_vs_torch': ok_max, 'timed_buffer_equals_fresh_call': ok_same}
    if rank == 0:
        from oracle import inference as oinf
        from oracle import triangulate as otri
        from pose_unsupervised_b200.utils import synth
        rigs = synth.camera_table(7, V, seed=0)
        flat_cams = [c for rig in rigs for c in rig]
        cams = [flat_cams[i] for i in index[:rows]]
        host = hm[:rows].cpu().numpy()
        ref_xy, ref_mv = oinf.get_final_preds(True, host, center[:rows], scale[:rows])
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref3d = otri.triangulate_poses(cams, chk.xy.cpu().numpy(), nviews=V)
        got = chk.poses3d.cpu().numpy()
        rel = float((np.abs(got - ref3d) / np.maximum(1.0, np.abs(ref3d))).max())
        info['xy_max_abs_diff_vs_oracle'] = float(np.abs(chk.xy.cpu().numpy() - ref_xy).max())
        info['poses_max_rel_diff_vs_oracle'] = rel
        info['oracle_ok'] = bool(rel < 1e-6 and info['xy_max_abs_diff_vs_oracle'] <= 1e-3
                                 and np.array_equal(chk.maxvals.cpu().numpy(), ref_mv[:, :, 0]))
    if world > 1:
        full = exch.gathered_poses(last_slot)
        lo = sum(exch.counts[:rank])
        info['gathered_equals_own_shard'] = bool(torch.equal(full[lo:lo + exch.counts[rank]], own_poses))
        info['gathered_frames'] = int(full.shape[0])
    info['ok'] = all(v for k, v in info.items() if isinstance(v, bool))
    return info


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
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    dynamic = args.decode_schedule != 'static'      # the library's default (profiles/r02_scaling.md)
    rt.set_decode_schedule(dynamic)
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
    # poses straight into the send buffer.  Four slots: the all-gather of step k-1 overlaps step k and is only
    # waited for when its slot comes round again, three steps later.
    nslots = 4 if world > 1 else 1
    exch = parallel.PoseExchange(nframes_total, J, dev, nslots=nslots)

    def compute(slot, fork=None, ev=None):
        aff = crop_affine(d_center, d_scale, (HW, HW), inv=1)
        if ev is not None:
            ev[0].record()
        res = lift_heatmaps(hm, None, None, table, nviews=V, post_process=True, affine=aff,
                            out_poses3d=exch.poses_view(slot), after_decode=fork)
        if ev is not None:
            ev[1].record()
        exch.stats_view(slot).zero_()
        mpjpe_stats(res.poses3d, gt, out=exch.stats_view(slot))
        return res

    def step(k, ev=None, join=True):
        return exch.pipelined_step(k, lambda slot, fork: compute(slot, fork if world > 1 else None, ev), join=join)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(warmup):
        step(k)
    fence()

    # CUDA graphs hold the steps (4 kernels and a memset each and -- forked onto the side stream after the
    # decode -- the NCCL all-gather of the other slot): at ~0.7 ms per step the Python/launch overhead of the
    # eager path is otherwise visible, above all at N > 1.  One graph spans GRAPH_STEPS steps so that only its
    # last all-gather is joined at the graph's end; the others are joined one step later, where their slot is
    # reused, and never wait.
    GRAPH_STEPS = 8 if world > 1 else 1
    graphs, graphed = {}, False

    def capture(nsteps):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            for k in range(nsteps):
                step(k, join=False)
            exch.join()
        return gph

    def run_steps(nsteps):
        """nsteps steps starting at slot 0 (nsteps is even or the last block): graph replays when captured."""
        if not graphed:
            for k in range(nsteps):
                step(k, join=False)
            exch.join()
            return
        full, rest = divmod(nsteps, GRAPH_STEPS)
        for _ in range(full):
            graphs[GRAPH_STEPS].replay()
        if rest:
            graphs[rest].replay()

    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for k in range(2 * nslots):
                    step(k)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graphs[GRAPH_STEPS] = capture(GRAPH_STEPS)
            if steps % GRAPH_STEPS:
                graphs[steps % GRAPH_STEPS] = capture(steps % GRAPH_STEPS)
            graphed = True
            run_steps(2 * GRAPH_STEPS)
        except Exception as e:   # fall back to eager launches, say so in the JSON line
            sys.stderr.write('CUDA graph capture failed (%s); timing eager launches\n' % e)
            graphs, graphed = {}, False
    fence()

    clocks = ClockSampler(local)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    fence()
    start.record()
    run_steps(steps)
    last_slot = (steps - 1) % GRAPH_STEPS % nslots if steps % GRAPH_STEPS else (GRAPH_STEPS - 1) % nslots
    if world > 1:
        exch.run(last_slot)                 # drain: the last step's exchange is inside the timed region
    stop.record()
    fence()
    clock_info = clocks.stop()
    ms_total = parallel.max_over_ranks(start.elapsed_time(stop), dev)
    value = nframes_total * steps / (ms_total * 1e-3)
    own_poses = exch.poses_view(last_slot).clone()
    # crop_affine_kernel, decode_tma_kernel, geometry_kernel<float,1>, mpjpe_kernel (+ a torch memset
    # of 4 doubles and, N > 1, ncclAllGather -- neither is counted: not ours)
    kernels_per_step = ['crop_affine_kernel', 'decode_tma_kernel', 'geometry_kernel<float,1>', 'mpjpe_kernel']

    aff = crop_affine(d_center, d_scale, (HW, HW), inv=1)
    verified = verify_lift(hm, pack, index, center, scale, aff, own_poses, exch, last_slot, rank, world)
    flag = torch.tensor([1.0 if verified['ok'] else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    verified['all_ranks_ok'] = bool(flag.item() == 1.0)

    # ---- the dominant kernel alone: CUDA events around every launch (eager pass, same work) ----
    ksteps = min(steps, 50)
    kernel_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(ksteps)]
    fence()
    for i in range(ksteps):
        compute(0, None, kernel_events[i])
    fence()
    lift_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    lift_ms = parallel.max_over_ranks(lift_ms, dev)
    # the lift is two kernels; the dominant one (decode_tma_kernel, the only pass over the
    # heatmaps) is timed alone through the decode entry point, same inputs, same launch
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

    # ---- end to end through the public numpy API, host buffers ------------------------------
    e2e_steps = max(1, min(3, steps))
    e2e_value = e2e_same = None
    h2d = d2h = 0
    e2e_pageable = None

    def e2e_run(host_hm):
        lift_heatmaps(host_hm, center, scale, table, nviews=V, post_process=True).numpy()   # warm-up
        fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = lift_heatmaps(host_hm, center, scale, table, nviews=V, post_process=True).numpy()
        torch.cuda.synchronize()
        dt = parallel.max_over_ranks(time.perf_counter() - t0, dev)
        return nframes_total * e2e_steps / dt, out

    if not args.no_e2e:
        pinned = torch.empty((B * V, J, HW, HW), dtype=torch.float32, pin_memory=True)
        pinned.copy_(hm)
        torch.cuda.synchronize()
        e2e_value, out = e2e_run(pinned.numpy())           # numpy view of pinned memory
        d2h = sum(a.nbytes for a in (out.xy, out.maxvals, out.poses3d, out.reproj_err))
        h2d = pinned.numel() * 4 + center.nbytes + scale.nbytes
        e2e_same = bool(np.array_equal(out.poses3d, own_poses.cpu().numpy()))
        if not args.no_pageable:
            pageable = np.empty(tuple(pinned.shape), dtype=np.float32)     # plain malloc'ed host memory
            np.copyto(pageable, pinned.numpy())
            pv, pout = e2e_run(pageable)
            e2e_pageable = {'value': pv, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                            'steps': e2e_steps, 'equals_pinned_result': bool(np.array_equal(pout.poses3d, out.poses3d)),
                            'note': 'pageable numpy input staged through two pinned buffers in 256 MiB chunks'}
            del pageable
        del pinned

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        achieved = kern_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            'metric': metric_name(), 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms_total / steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 decode, f64 lift', 'data': 'synthetic',
            'config': dict(workload_config(B), decode_schedule='dynamic' if dynamic else 'static', cuda_graph=graphed, steps_per_graph=GRAPH_STEPS if graphed else None,
                           exchange='none (1 GPU)' if world == 1 else
                           'one ncclAllGather of [poses | MPJPE sums] per step, double-buffered: the gather of '
                           'step k-1 is forked onto a side stream after the decode of step k, runs under its lift / MPJPE '
                           'kernels and is joined one step later; the last one is drained inside the timed region'),
            'clocks': clock_info,
            'e2e': None if args.no_e2e else {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps, 'host_memory': 'pinned',
                    'equals_device_resident_result': e2e_same,
                    'api': 'pose_unsupervised_b200.multiviews.triangulate.lift_heatmaps (numpy in, numpy out)'},
            'e2e_pageable': e2e_pageable,
            'gpu_launches': steps * len(kernels_per_step),
            'kernels_per_step': kernels_per_step,
            'verified': verified['ok'] and verified['all_ranks_ok'],
            'verification': verified,
            'roofline': {'bound': 'hbm', 'kernel': kern_name, 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None, 'traffic_source': None,
                         'peak_source': peak_src, 'kernel_ms': kern_ms,
                         'algorithmic_bytes_per_launch': kern_bytes,
                         'lift_path_ms': lift_ms,
                         'whole_path_frac': B * bytes_per_frame() / (lift_ms * 1e-3) / 1e9 / peak},
            'cpu_baseline': None,
        }
        traffic = os.path.join(ROOT, 'profiles', 'decode_tma_traffic.json')
        if os.path.exists(traffic):
            try:
                with open(traffic) as f:
                    t = json.load(f)
                if t.get('frames_per_launch') == B and t.get('kernel') == kern_name and \
                        (t.get('views'), t.get('joints'), t.get('hw')) == (V, J, HW):
                    line['roofline']['traffic'] = t.get('dram_bytes_per_launch')
                    line['roofline']['traffic_source'] = 'profiles/decode_tma_traffic.json: %s' % t.get('source')
            except Exception:
                pass

    # drop everything that captured a collective before the group goes away (NCCL's teardown waits
    # for CUDA graphs that hold its kernels)
    graphs = None
    gc.collect()
    torch.cuda.synchronize()

    if rank == 0 and world == 1 and not args.no_secondary:
        del hm
        torch.cuda.empty_cache()
        line['secondary'] = secondary_legs(args)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_leg()
        print(json.dumps(line))
        sys.stdout.flush()
    parallel.shutdown()


# ---------------------------------------------------------------------------------------
# secondary workloads: configs[2] RPSM, configs[3] pseudo-label pass, configs[4] sweep
# ---------------------------------------------------------------------------------------
def rpsm_problem(B, base=8):
    """B frames of the RPSM workload (4 views x 17 joints, Gaussian heatmaps of `base` distinct poses)."""
    import types
    import torch
    from pose_unsupervised_b200.multiviews import pictorial
    from pose_unsupervised_b200.multiviews.body import HumanBody
    from pose_unsupervised_b200.multiviews.cameras import CameraTable
    from pose_unsupervised_b200.utils import synth
    body = HumanBody.h36m17()
    edges = body.edges()
    cfg = types.SimpleNamespace(
        NETWORK=types.SimpleNamespace(IMAGE_SIZE=np.array([256, 256]), HEATMAP_SIZE=np.array([64, 64])),
        PICT_STRUCT=types.SimpleNamespace(FIRST_NBINS=16, RECUR_NBINS=2, RECUR_DEPTH=10, GRID_SIZE=2000,
                                          LIMB_LENGTH_TOLERANCE=150))
    poses = synth.random_poses(base, seed=1)
    avg = {e: float(np.mean([np.linalg.norm(p[e[0]] - p[e[1]]) for p in synth.random_poses(64, seed=99)]))
           for e in edges}
    table = pictorial.PairwiseTable.from_limb_lengths(avg, body, 2000, 16)
    hms, cams, boxes_all, centers, scales, roots, limbs, limb_dicts = [], [], [], [], [], [], [], []
    for f in range(base):
        cam = synth.camera_ring(4, seed=50 + f)
        boxes = synth.crop_box(cam, poses[f])
        hms.append(synth.gaussian_heatmaps(cam, boxes, poses[f], 64, 256, 2.0, 0.02, seed=f))
        cams.append(cam)
        boxes_all.append(boxes)
        roots.append(poses[f][0] + [20.0, -30.0, 10.0])
        centers.append([b['center'] for b in boxes])
        scales.append([b['scale'] for b in boxes])
        limb = synth.limb_lengths(poses[f], edges)
        limb_dicts.append(limb)
        limbs.append([limb[e] for e in edges])
    rep = [i % base for i in range(B)]
    hm = torch.from_numpy(np.array(hms)).cuda()[rep].contiguous()
    ctab = CameraTable.from_cameras([c for i in rep for c in cams[i]])
    cen = np.array([centers[i] for i in rep]).reshape(-1, 2)
    sca = np.array([scales[i] for i in rep]).reshape(-1, 2)
    roo = np.array([roots[i] for i in rep])
    lim = np.array([limbs[i] for i in rep])
    host = dict(hms=hms, cams=cams, boxes=boxes_all, roots=roots, limb_dicts=limb_dicts, avg=avg, poses=poses)
    return dict(args=(ctab, hm, cen, sca, roo, lim, table, cfg, body), cfg=cfg, host=host, base=base)


def rpsm_leg(B=592, steps=5, cpu_frames=2):
    """configs[2]: frames/s of pb200_rpsm, checked against and timed beside the oracle port of
    lib/multiviews/pictorial.py:rpsm on `cpu_frames` of the same frames (one host core, like the
    reference's single-process test_rpsm.py)."""
    import torch
    from pose_unsupervised_b200.multiviews import pictorial
    prob = rpsm_problem(B)
    for _ in range(2):
        out = pictorial.rpsm_batch(*prob['args'])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        out = pictorial.rpsm_batch(*prob['args'])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    got = out.cpu().numpy()
    h = prob['host']
    err = float(np.mean(np.linalg.norm(got[:prob['base']] - h['poses'], axis=2)))
    res = {'workload': 'configs[2]: RPSM 4 views x 17 joints, 16^3 then 10 x 2^3', 'frames': B,
           'ms_per_step': ms, 'frames_per_s': B / (ms * 1e-3), 'mpjpe_mm_vs_synthetic_gt': err,
           'gpu_launches': steps, 'kernel': 'rpsm_onchip_kernel'}
    if cpu_frames > 0:
        from oracle import pictorial as opict
        from oracle.body import h36m17
        obody = h36m17()
        opw = opict.level0_pairwise(2000, h['avg'], 16, obody)
        t0 = time.perf_counter()
        same = True
        for f in range(cpu_frames):
            ref = opict.rpsm(h['cams'][f], h['hms'][f], h['boxes'][f], h['roots'][f], h['limb_dicts'][f], opw,
                             prob['cfg'], obody)
            same = same and bool(np.array_equal(ref, got[f]))
        dt = time.perf_counter() - t0
        res['cpu_baseline'] = {'value': cpu_frames / dt, 'unit': 'frames/s', 'cores': 1, 'kind': 'port',
                               'sample': '%d frames of the same batch, oracle port of pictorial.rpsm' % cpu_frames}
        res['poses_equal_oracle'] = same
        res['speedup_vs_cpu_port'] = res['frames_per_s'] / res['cpu_baseline']['value']
    return res


def device_observations(B, nviews, dev, seed):
    """Synthetic 2D input of the pseudo-label pass, generated on the device with torch (input
    rendering only): poses -> plumb-bob projections through the frame's rig + 2 px noise + 10 %
    outliers at 50 px, confidences U(0.04, 1.12)."""
    import torch
    from pose_unsupervised_b200.multiviews.cameras import pack_camera
    from pose_unsupervised_b200.utils import synth
    g = torch.Generator(device=dev).manual_seed(seed)
    rigs = synth.camera_table(7, nviews, seed=0)
    pack = np.array([pack_camera(c) for rig in rigs for c in rig])
    P = torch.from_numpy(pack).to(dev)                                              # [7*V, 24]
    subj = torch.randint(0, 7, (B,), generator=g, device=dev)
    base = torch.from_numpy(synth.random_poses(1024, seed=5)).to(dev)
    poses = base[torch.randint(0, 1024, (B,), generator=g, device=dev)] + \
        15.0 * torch.randn((B, 17, 3), generator=g, device=dev, dtype=torch.float64)
    cam = P[(subj[:, None] * nviews + torch.arange(nviews, device=dev)[None])]        # [B, V, 24]
    R, T = cam[..., :9].view(B, nviews, 3, 3), cam[..., 9:12]
    xc = torch.einsum('bvrc,bvjc->bvjr', R, poses[:, None] - T[:, :, None])
    x, y = xc[..., 0] / xc[..., 2], xc[..., 1] / xc[..., 2]
    k, p = cam[..., 16:19], cam[..., 19:21]
    r2 = x * x + y * y
    barrel = 1 + k[..., 0:1] * r2 + k[..., 1:2] * r2 ** 2 + k[..., 2:3] * r2 ** 3
    xd = x * barrel + 2 * p[..., 0:1] * x * y + p[..., 1:2] * (r2 + 2 * x * x)
    yd = y * barrel + p[..., 0:1] * (r2 + 2 * y * y) + 2 * p[..., 1:2] * x * y
    obs = torch.stack([cam[..., 12:13] * xd + cam[..., 14:15], cam[..., 13:14] * yd + cam[..., 15:16]], dim=-1)
    obs = obs + 2.0 * torch.randn(obs.shape, generator=g, device=dev, dtype=torch.float64)
    bad = torch.rand(obs.shape[:-1], generator=g, device=dev) < 0.10
    obs = obs + bad[..., None] * 50.0 * torch.randn(obs.shape, generator=g, device=dev, dtype=torch.float64)
    conf = 0.04 + 1.08 * torch.rand((B * nviews, 17), generator=g, device=dev)
    index = (subj[:, None] * nviews + torch.arange(nviews, device=dev)[None]).reshape(-1).to(torch.int32)
    return rigs, pack, subj, index, poses, obs.reshape(B * nviews, 17, 2).to(torch.float32).contiguous(), conf


def pseudo_leg(total, rank, world, dev, steps=5, parts=True):
    """configs[3](ii): the pseudo-label pass of run/test/test_pseudo_label.py from 2D locations."""
    import types
    import torch
    from pose_unsupervised_b200 import parallel
    from pose_unsupervised_b200.core.loss import FundamentalTable, epipolar_residuals
    from pose_unsupervised_b200.multiviews.cameras import CameraTable
    from pose_unsupervised_b200.multiviews.triangulate import ransac, reproject_poses
    lo, hi = parallel.frame_shard(total, rank, world)
    B = hi - lo
    rigs, pack, subj, index, poses, d_obs, d_conf = device_observations(B, 4, dev, 100 + rank)
    table = CameraTable(torch.from_numpy(pack).to(dev), index)
    ftab = FundamentalTable.from_cameras({s_: rigs[s_] for s_ in range(7)})
    cfg = types.SimpleNamespace(DATASET=types.SimpleNamespace(NO_DISTORTION=False),
                                PSEUDO_LABEL=types.SimpleNamespace(REPROJ_THRE=10.0, NUM_INLIERS=3))
    slots = ftab.slots(subj.cpu().numpy())
    exch = parallel.PoseExchange(total, 17, dev, nslots=1) if world > 1 else None

    def step():
        vis = d_conf > 0.7                                            # test_pseudo_label.py:194
        vis = ransac(d_obs, table, vis, cfg)                          # :221
        proj, pvis, pts = reproject_poses(d_obs, table, vis, False, return_points=True)   # :237
        resid = epipolar_residuals(proj, slots, ftab)                 # test_fund_mtx.py:56-69 on the labels
        if exch is not None:
            exch.poses_view().copy_(pts)
            exch.run()
        return proj, pvis, pts, resid

    def timed(fn, n):
        for _ in range(2):
            out = fn()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            out = fn()
        b.record()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(a.elapsed_time(b) / n, dev), out

    ms, out = timed(step, steps)
    proj, pvis, pts, resid = out
    keep = float(pvis.float().mean())
    lifted = pvis.view(B, 4, 17)[:, 0] > 0
    err = float((pts - poses)[lifted].abs().mean()) if B else 0.0
    per_frame = 4 * 17 * 12 + 32 + 4 * 17 * 8 + 4 * 17 + 17 * 24 + 12 * 17 * 8
    res = {'workload': 'configs[3](ii): pseudo-label pass from 2D locations (conf>0.7, RANSAC 3 inliers/10 px, '
                       'reproject, epipolar residuals)', 'frames': total, 'n_gpus': world, 'ms_per_step': ms,
           'frames_per_s': total / (ms * 1e-3), 'algorithmic_GBps': total * per_frame / (ms * 1e-3) / 1e9,
           'labels_kept': keep, 'mean_abs_3d_err_mm': err, 'gpu_launches': 3 * steps,
           'kernels_per_step': ['ransac_compact_kernel', 'geometry_kernel<float,1>', 'epipolar_kernel']}
    if parts:
        vis0 = d_conf > 0.7
        all_vis = torch.ones_like(vis0)
        vis1 = ransac(d_obs, table, vis0, cfg)
        res['parts_ms'] = {
            'ransac_conf_gt_0.7': timed(lambda: ransac(d_obs, table, vis0, cfg), steps)[0],
            'ransac_all_visible': timed(lambda: ransac(d_obs, table, all_vis, cfg), steps)[0],
            'reproject': timed(lambda: reproject_poses(d_obs, table, vis1, False, return_points=True), steps)[0],
            'epipolar': timed(lambda: epipolar_residuals(proj, slots, ftab), steps)[0]}
    return res


def sweep_leg(frames=4096, steps=20, cpu=False):
    """configs[4] on one GPU: decode + lift for 2/4/8 views x 64^2/96^2 (frames/s, fraction of the HBM
    roofline of SURVEY.md section 8d); with cpu=True the CPU port is timed beside each shape."""
    import torch
    from pose_unsupervised_b200 import runtime as rt
    from pose_unsupervised_b200.multiviews.cameras import CameraTable
    peak, _ = measured_hbm_peak()
    rows = []
    dev = torch.device('cuda', torch.cuda.current_device())
    for hw in (64, 96):
        for v in (2, 4, 8):
            rigs, subj, pack, index, center, scale = make_side_inputs(frames, seed=7, nviews=v)
            table = CameraTable.from_arrays(pack, index)
            g = torch.Generator(device=dev).manual_seed(99)
            hm = torch.rand((frames * v, 17, hw, hw), generator=g, device=dev, dtype=torch.float32)
            ms = time_lift_device(hm, rt.to_device(center), rt.to_device(scale), table, frames, v, steps)
            bpf = bytes_per_frame(v, 17, hw)
            row = {'views': v, 'hw': hw, 'frames': frames, 'ms_per_step': ms, 'frames_per_s': frames / (ms * 1e-3),
                   'bytes_per_frame': bpf, 'roofline_frac': frames * bpf / (ms * 1e-3) / 1e9 / peak}
            if cpu:
                row['cpu_baseline'] = cpu_baseline_leg(shape=(v, 17, hw), seconds=3.0)
            rows.append(row)
            del hm
            torch.cuda.empty_cache()
    return rows


def secondary_legs(args):
    import torch
    out = {}
    dev = torch.device('cuda', torch.cuda.current_device())
    for name, fn in (('rpsm', lambda: rpsm_leg()),
                     ('pseudo_label', lambda: pseudo_leg(1000000, 0, 1, dev)),
                     ('sweep', lambda: sweep_leg())):
        try:
            out[name] = fn()
        except Exception as e:            # the headline line must survive; the failure is reported, not hidden
            out[name] = {'error': '%s: %s' % (type(e).__name__, e)}
        torch.cuda.empty_cache()
    return out


def run_rpsm(args):
    import torch
    torch.cuda.set_device(0)
    B = args.frames if args.frames != 4096 else 592
    print(json.dumps(rpsm_leg(B, steps=args.steps if args.steps else 5, cpu_frames=0 if args.no_cpu_baseline else 2)))


def init_ranks():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    return world, rank, dev


def run_pseudo(args):
    """configs[3](ii) on `--frames` frames in total (default 1 M), sharded by frame over the ranks."""
    from pose_unsupervised_b200 import parallel
    world, rank, dev = init_ranks()
    total = args.frames if args.frames != 4096 else 1000000
    res = pseudo_leg(total, rank, world, dev, steps=args.steps if args.steps else 5, parts=(world == 1))
    if rank == 0:
        print(json.dumps(res))
    parallel.shutdown()


def run_pseudo_hm(args):
    """configs[3](i): the pseudo-label pass FROM HEATMAPS -- `--frames` frames in total (default 1 M),
    sharded over the ranks and streamed through lift_heatmaps(fundamental=...) in device-generated
    chunks (decode + conf threshold + triangulate + reproject + epipolar residuals per chunk).  The
    synthetic chunks are generated before the timed region; a rank keeps `--resident-chunks`
    distinct chunks in HBM and cycles through them, so every step streams fresh bytes from HBM
    (each chunk is far larger than L2)."""
    import torch
    from pose_unsupervised_b200 import parallel, runtime as rt
    from pose_unsupervised_b200.core.loss import FundamentalTable
    from pose_unsupervised_b200.multiviews.cameras import CameraTable
    from pose_unsupervised_b200.multiviews.triangulate import lift_heatmaps
    from pose_unsupervised_b200.utils.transforms import crop_affine
    world, rank, dev = init_ranks()
    total = args.frames if args.frames != 4096 else 1000000
    lo, hi = parallel.frame_shard(total, rank, world)
    mine = hi - lo
    chunk = args.chunk_frames
    resident = min(args.resident_chunks, (mine + chunk - 1) // chunk)
    rigs, subj, pack, index, center, scale = make_side_inputs(chunk, seed=rank)
    table = CameraTable.from_arrays(pack, index)
    ftab = FundamentalTable.from_cameras({s_: rigs[s_] for s_ in range(7)})
    slots = ftab.slots(subj)
    aff = crop_affine(rt.to_device(center), rt.to_device(scale), (HW, HW), inv=1)
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    chunks = [torch.rand((chunk * V, J, HW, HW), generator=g, device=dev, dtype=torch.float32) for _ in range(resident)]
    poses = torch.empty((mine, J, 3), dtype=torch.float64, device=dev)
    nchunks = (mine + chunk - 1) // chunk

    def one_pass():
        for c in range(nchunks):
            n = min(chunk, mine - c * chunk)
            hm = chunks[c % resident]
            lift_heatmaps(hm[:n * V], None, None, CameraTable(table.pack, table.index[:n * V]), nviews=V,
                          post_process=True, conf_thre=0.7, affine=aff[:n * V],
                          out_poses3d=poses[c * chunk:c * chunk + n], fundamental=ftab, subjects=slots[:n])
        if world > 1:
            parallel.gather_poses(poses, total)

    one_pass()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    steps = args.steps if args.steps else 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        one_pass()
    b.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(a.elapsed_time(b) / steps, dev)
    if rank == 0:
        peak, _ = measured_hbm_peak()
        print(json.dumps({'workload': 'configs[3](i): pseudo-label pass from heatmaps (decode, conf>0.7, triangulate, '
                                      'reproject, epipolar residuals), chunked', 'frames': total, 'n_gpus': world,
                          'chunk_frames': chunk, 'chunks_per_rank': nchunks, 'resident_chunks': resident,
                          'ms_per_pass': ms, 'frames_per_s': total / (ms * 1e-3),
                          'roofline_frac': total * bytes_per_frame() / (ms * 1e-3) / 1e9 / (peak * world),
                          'gpu_launches': steps * nchunks * 2}))
    parallel.shutdown()


def run_sweep(args):
    import torch
    torch.cuda.set_device(0)
    print(json.dumps({'workload': 'configs[4]: view/resolution sweep, 1 GPU, decode + lift',
                      'rows': sweep_leg(args.frames, args.steps if args.steps else 20, cpu=not args.no_cpu_baseline)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=0)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='lift', choices=['lift', 'rpsm', 'pseudo', 'pseudo-hm', 'sweep'])
    ap.add_argument('--views', type=int, default=4)
    ap.add_argument('--joints', type=int, default=17)
    ap.add_argument('--hw', type=int, default=64, help='heatmap side')
    ap.add_argument('--frames', type=int, default=4096, help='frames per GPU (lift) / in total (pseudo, pseudo-hm)')
    ap.add_argument('--chunk-frames', type=int, default=8192, help='pseudo-hm: frames per device chunk')
    ap.add_argument('--resident-chunks', type=int, default=4, help='pseudo-hm: distinct chunks kept in HBM per rank')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the RPSM / pseudo-label / sweep legs of the default run')
    ap.add_argument('--no-pageable', action='store_true', help='skip the pageable-memory e2e leg')
    ap.add_argument('--no-e2e', action='store_true', help='sweep runs only: skip the host-buffer legs (e2e is then null)')
    ap.add_argument('--decode-schedule', default='auto', choices=['auto', 'static', 'dynamic'],
                    help='auto = the library default (dynamic: strided share + claimed tail); static for A/B runs')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph replay')
    args = ap.parse_args()
    global V, J, HW
    V, J, HW = args.views, args.joints, args.hw
    if args.impl == 'reference':
        run_reference_arm(args)
    elif args.workload == 'rpsm':
        run_rpsm(args)
    elif args.workload == 'pseudo':
        run_pseudo(args)
    elif args.workload == 'pseudo-hm':
        run_pseudo_hm(args)
    elif args.workload == 'sweep':
        run_sweep(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
