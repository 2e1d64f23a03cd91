"""CPU test of the N>1 exchange step (gloo, world size 2): all-gather of per-frame poses
and all-reduce of the MPJPE partial sums, with an uneven frame split."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pose_unsupervised_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nframes, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        pred = torch.from_numpy(rng.normal(0, 100, (nframes, 17, 3)))
        gt = torch.from_numpy(rng.normal(0, 100, (nframes, 17, 3)))
        lo, hi = parallel.frame_shard(nframes, rank, world)
        full = parallel.gather_poses(pred[lo:hi].clone(), nframes)
        err = (pred[lo:hi] - gt[lo:hi]).norm(dim=2)
        stats = torch.stack([err.sum(), (err * err).sum(), err.max(),
                             torch.tensor(float(err.numel()), dtype=torch.float64)])
        red = parallel.reduce_mpjpe(stats)
        t = parallel.max_over_ranks(1.0 + rank, torch.device('cpu'))
        # the one-collective form used by bench.py
        ex = parallel.PoseExchange(nframes, 17, torch.device('cpu'))
        ex.poses_view().copy_(pred[lo:hi])
        ex.stats_view().copy_(stats)
        ex.run()
        red2 = ex.reduced_stats()
        assert torch.equal(ex.gathered_poses(), pred) and torch.allclose(red2, red)
        # double-buffered form: the gather of step k-1 is issued in step k, the last one drained
        ex2 = parallel.PoseExchange(nframes, 17, torch.device('cpu'), nslots=2)
        for k in range(3):
            def compute(slot, fork, k=k):
                ex2.poses_view(slot).copy_(pred[lo:hi] + k)
                if k == 1:
                    fork()                      # mid-step fork, as bench.py does after the decode
                ex2.stats_view(slot).copy_(stats)
            ex2.pipelined_step(k, compute)
            if k >= 1:
                assert torch.equal(ex2.gathered_poses((k - 1) % 2), pred + (k - 1))
        ex2.run(2 % 2)
        assert torch.equal(ex2.gathered_poses(0), pred + 2) and torch.allclose(ex2.reduced_stats(0), red)
        all_err = (pred - gt).norm(dim=2)
        ok = torch.equal(full, pred) and abs(float(red[0]) - float(all_err.sum())) < 1e-6 \
            and float(red[2]) == float(all_err.max()) and float(red[3]) == all_err.numel() \
            and t == float(world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run(nframes):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nframes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_even_split():
    _run(8)


def test_uneven_split():
    _run(7)
