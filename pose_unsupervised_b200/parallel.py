"""Frame-sharded multi-GPU driver: one process per GPU, NCCL over NVLink.

Frames are independent in every stage of the lifting path, so the
data path has no collective: rank r owns the contiguous frame block
[r*B/R, (r+1)*B/R) of the view-minor arrays.  The only exchange step comes after
the compute: an all-gather of the per-frame 3D poses and an all-reduce of the MPJPE
partial sums (run/test/test_triangulate.py:98-101), both issued on the compute
stream.  The same code runs on CPU tensors with the gloo backend (tests).
"""
import torch
import torch.distributed as dist


def frame_shard(nframes, rank, world):
    """Contiguous frame range of ``rank``; the remainder goes to the last rank."""
    per = nframes // world
    lo = rank * per
    hi = nframes if rank == world - 1 else lo + per
    return lo, hi


def shard_rows(array, nviews, rank, world):
    """Rows [B*V, ...] view-minor -> the rows of this rank's frames."""
    nframes = array.shape[0] // nviews
    lo, hi = frame_shard(nframes, rank, world)
    return array[lo * nviews:hi * nviews]


def gather_poses(local_poses, nframes, group=None):
    """All-gather of per-frame 3D poses: local [B_r, J, 3] -> [nframes, J, 3] on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_poses
    sizes = [frame_shard(nframes, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    tail = tuple(local_poses.shape[1:])
    if len(set(counts)) == 1:
        out = local_poses.new_empty((nframes,) + tail)
        dist.all_gather_into_tensor(out, local_poses.contiguous(), group=group)
        return out
    # uneven split (the last rank takes the remainder): pad to the largest shard so that one
    # fixed-size all-gather suffices, then drop the padding
    cap = max(counts)
    padded = local_poses.new_zeros((cap,) + tail)
    padded[:local_poses.shape[0]] = local_poses
    gathered = local_poses.new_empty((world * cap,) + tail)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    gathered = gathered.reshape((world, cap) + tail)
    return torch.cat([gathered[r, :counts[r]] for r in range(world)], dim=0)


def reduce_mpjpe(stats, group=None):
    """All-reduce of [sum, sumsq, max, count]: SUM for 0,1,3 and MAX for 2."""
    if dist.get_world_size(group) == 1:
        return stats
    sums = stats[[0, 1, 3]].clone()
    mx = stats[2:3].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return torch.stack([sums[0], sums[1], mx[0], sums[2]])


class PoseExchange(object):
    """The exchange step as ONE collective: every rank contributes [its poses | its 4 MPJPE sums].

    The send buffer is where the lift kernel writes its 3D poses (``poses_view``) and where
    ``mpjpe_stats`` accumulates (``stats_view``), so nothing is copied before the all-gather;
    buffers are allocated once, which also makes the whole step capturable in a CUDA graph.
    Shards may be uneven (the last rank takes the remainder): every rank sends ``cap`` frames.

    ``nslots`` > 1 gives independent send/receive buffers: with two slots the all-gather of step
    k-1 runs on a side stream underneath the kernels of step k (``pipelined_step``), so the
    collective's latency leaves the critical path.
    """

    def __init__(self, nframes, njoints, device, group=None, dtype=torch.float64, nslots=1):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nframes, self.njoints, self.nslots = nframes, njoints, nslots
        self.counts = [hi - lo for lo, hi in (frame_shard(nframes, r, self.world) for r in range(self.world))]
        self.cap = max(self.counts)
        self.width = self.cap * njoints * 3 + 4
        self.send = [torch.zeros(self.width, dtype=dtype, device=device) for _ in range(nslots)]
        self.recv = [torch.zeros(self.world * self.width, dtype=dtype, device=device) for _ in range(nslots)]
        self.side = torch.cuda.Stream(device=device) if (device.type == 'cuda' and nslots > 1) else None
        self._pending = False     # a collective forked onto the side stream has not been joined yet
        self._gathered = [None] * nslots   # per slot: event of the un-joined collective that reads it

    def poses_view(self, slot=0):
        b = self.counts[self.rank]
        return self.send[slot][:b * self.njoints * 3].view(b, self.njoints, 3)

    def stats_view(self, slot=0):
        return self.send[slot][self.width - 4:]

    def run(self, slot=0):
        """All-gather of slot ``slot`` on the current stream."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group)
        else:
            self.recv[slot].copy_(self.send[slot])

    def pipelined_step(self, k, compute, join=True):
        """Step ``k`` of a multi-buffered loop: ``compute(slot, fork)`` fills slot ``k % nslots`` on the
        current stream; the all-gather of the previous step's slot runs on the side stream from the
        moment ``compute`` calls ``fork()`` (at the latest when it returns).  Forking after the
        HBM-bound decode and before the small latency-bound kernels puts the collective where SMs are
        idle.  With ``join=True`` the current stream waits for the collective at the end of the step.
        With ``join=False`` the wait moves to the ``fork()`` of the step that REUSES the slot the
        collective reads (``nslots - 1`` steps later), so ranks may drift apart by that many steps before
        anybody waits; call ``join()`` after the last step of a block (required before the end of a
        CUDA-graph capture).  Fork and join are stream / event waits, so any number of steps is
        capturable in one CUDA graph.  The last step's slot is still to be gathered afterwards:
        ``run((K-1) % nslots)``.
        """
        cur, prev = k % self.nslots, (k - 1) % self.nslots
        if self.world == 1:
            return compute(cur, lambda: None)
        state = {'forked': False}
        main = torch.cuda.current_stream() if self.side is not None else None

        def fork():
            if state['forked']:
                return
            state['forked'] = True
            if self.side is None:                   # CPU tensors (gloo tests): same order, no overlap
                self.run(prev)
                return
            done = self._gathered[cur]
            if done is not None:
                main.wait_event(done)               # the collective that read slot `cur` has finished
                self._gathered[cur] = None
            self.side.wait_stream(main)             # everything issued so far, step k-1 included, is ahead
            with torch.cuda.stream(self.side):
                self.run(prev)
                ev = torch.cuda.Event()
                ev.record(self.side)
            self._gathered[prev] = ev
            self._pending = True

        out = compute(cur, fork)
        fork()
        if join:
            self.join()
        return out

    def join(self):
        """The current stream waits for every collective in flight on the side stream."""
        if self.side is not None and self.world > 1 and self._pending:
            torch.cuda.current_stream().wait_stream(self.side)
        self._pending = False
        self._gathered = [None] * self.nslots

    def gathered_poses(self, slot=0):
        rows = self.recv[slot].view(self.world, self.width)
        return torch.cat([rows[r, :self.counts[r] * self.njoints * 3].view(self.counts[r], self.njoints, 3)
                          for r in range(self.world)], dim=0)

    def reduced_stats(self, slot=0):
        st = self.recv[slot].view(self.world, self.width)[:, self.width - 4:]
        return torch.stack([st[:, 0].sum(), st[:, 1].sum(), st[:, 2].max(), st[:, 3].sum()])


def shutdown(timeout_s=20.0):
    """Leave the process group without hanging: callers drop every CUDA graph that captured a
    collective first (NCCL's communicator teardown waits for them), then this barriers, syncs and
    destroys the group under a watchdog that ends the process if teardown does not return."""
    import gc
    import os
    import sys
    import threading
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if not dist.is_initialized():
        return
    dog = threading.Timer(timeout_s, lambda: os._exit(0))
    dog.daemon = True
    dog.start()
    try:
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.destroy_process_group()
    finally:
        dog.cancel()


def max_over_ranks(value, device, group=None):
    """Scalar max over ranks (timing: a step takes as long as its slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
