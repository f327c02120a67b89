"""Frame-batch sharding across the GPUs of one box.

Frames are independent units (every gs_* op takes one image), so the only multi-GPU data
movement the path has is one scatter of uint8 frames from the root rank and one gather of the
results -- grouped ncclSend/ncclRecv over NVLink 5 / NVSwitch in production (torch.distributed's
batch_isend_irecv = ncclGroupStart ... ncclGroupEnd), gloo in the CPU tests.  One process per GPU;
nothing else crosses ranks.  The root's NVLink egress (900 GB/s per direction) bounds the scatter,
its ingress the gather.

    shard_range       contiguous frame range of a rank
    scatter_frames    root (n, ...) -> every rank its shard                (blocking on the stream)
    gather_frames     inverse, one tensor
    gather_many       inverse, several result tensors in ONE group
    ShardedRun        scatter -> per-rank pipeline -> gather, whole-shard or in chunks whose
                      transfers overlap the neighbouring chunks' compute (side stream)
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous frame range [lo, hi) of `rank`: GPU g gets frames [g*n/G, (g+1)*n/G)"""
    return (n * rank) // world, (n * (rank + 1)) // world


def _run(ops):
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def scatter_frames(frames, n, frame_shape, dtype, device, src=0, group=None, out=None):
    """Root holds `frames` (n, *frame_shape); every rank returns its contiguous shard (into `out` if given).
    Grouped point-to-point sends (NCCL has no native scatter; shards may be ragged)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n, rank, world)
    mine = out if out is not None else torch.empty((hi - lo,) + tuple(frame_shape), dtype=dtype, device=device)
    assert mine.shape[0] == hi - lo
    if world == 1:
        mine.copy_(frames[lo:hi])
        return mine
    ops = []
    if rank == src:
        for r in range(world):
            a, b = shard_range(n, r, world)
            if r == src:
                mine.copy_(frames[a:b])
            elif b > a:
                ops.append(dist.P2POp(dist.isend, frames[a:b], r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, mine, src, group))
    _run(ops)
    return mine


def gather_many(mine, n, dst=0, group=None, out=None):
    """mine: list of this rank's result tensors (shard rows first); root returns the list assembled in frame
    order (into `out` if given: preallocated (n, ...) tensors), the others None.  One NCCL group for all."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        if out is None:
            return list(mine)
        for o, m in zip(out, mine):
            o.copy_(m)
        return out
    ops = []
    res = None
    if rank == dst:
        res = out if out is not None else [torch.empty((n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device) for m in mine]
        for r in range(world):
            a, b = shard_range(n, r, world)
            for o, m in zip(res, mine):
                if r == dst:
                    o[a:b].copy_(m)
                elif b > a:
                    ops.append(dist.P2POp(dist.irecv, o[a:b], r, group))
    elif mine[0].shape[0] > 0:
        for m in mine:
            ops.append(dist.P2POp(dist.isend, m.contiguous(), dst, group))
    _run(ops)
    return res


def gather_frames(mine, n, dst=0, group=None):
    """Inverse of scatter_frames for one tensor: root returns (n, ...) in frame order, others None."""
    if dist.get_world_size(group) == 1:
        return mine
    r = gather_many([mine], n, dst, group)
    return r[0] if r is not None else None


class ShardedRun:
    """One scatter / one gather around a per-rank frame pipeline.

    root_frames : (n_total, h, w) uint8 on the root rank's device (ignored elsewhere)
    pipe        : object with .run(frames, lo) writing rows [lo, lo+m) of its result buffers and
                  .results(m) -> dict of tensors (grayskull_b200.pipeline.FramePipeline)
    keys        : which results travel back to the root
    Timing (CUDA events on this rank, ms): scatter, compute, gather, total -- callers reduce with MAX over ranks.
    """

    def __init__(self, pipe, n_total, h, w, device, keys=("sobel", "kps", "kcounts", "rects", "rcounts"), src=0, group=None):
        self.pipe, self.n_total, self.h, self.w, self.device, self.keys, self.src, self.group = pipe, n_total, h, w, device, keys, src, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        self.m = self.hi - self.lo
        self.mine = torch.empty((self.m, h, w), dtype=torch.uint8, device=device)
        self.gathered = None
        if self.rank == src:
            res = pipe.results(0)
            self.gathered = [torch.empty((n_total,) + tuple(res[k].shape[1:]), dtype=res[k].dtype, device=device) for k in keys]
        self.cuda = torch.device(device).type == "cuda"
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None      # CPU / gloo: same schedule, no streams

    def bytes_scattered(self):
        """bytes leaving the root (its own shard is a local copy)"""
        lo, hi = shard_range(self.n_total, self.src, self.world)
        return (self.n_total - (hi - lo)) * self.h * self.w

    def bytes_gathered(self):
        lo, hi = shard_range(self.n_total, self.src, self.world)
        res = self.pipe.results(0)
        per_frame = 0
        for k in self.keys:
            e = res[k].element_size()
            for d in res[k].shape[1:]:
                e *= int(d)
            per_frame += e
        return (self.n_total - (hi - lo)) * per_frame

    # ---- whole shard: scatter, compute, gather back to back on the current stream --------------------------
    def run_serial(self, root_frames):
        if not self.cuda:
            scatter_frames(root_frames, self.n_total, (self.h, self.w), torch.uint8, self.device, self.src, self.group, out=self.mine)
            self.pipe.run(self.mine, 0)
            res = self.pipe.results(self.m)
            gather_many([res[k] for k in self.keys], self.n_total, self.src, self.group, out=self.gathered)
            return None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        scatter_frames(root_frames, self.n_total, (self.h, self.w), torch.uint8, self.device, self.src, self.group, out=self.mine)
        ev[1].record()
        self.pipe.run(self.mine, 0)
        ev[2].record()
        res = self.pipe.results(self.m)
        gather_many([res[k] for k in self.keys], self.n_total, self.src, self.group, out=self.gathered)
        ev[3].record()
        return ev

    # ---- chunks: chunk c's transfers ride a side stream while chunk c-1 / c+1 compute ------------------------
    def run_overlapped(self, root_frames, nchunks):
        """Every rank's shard is cut into `nchunks` equal pieces (the last may be short); piece c of ALL ranks is
        scattered in one NCCL group, so the root's egress stays busy while piece c-1 is being processed, and piece
        c's results are gathered while piece c+1 is processed.  On a CPU group (gloo tests) the same schedule runs
        without streams."""
        import contextlib
        cuda = self.cuda
        side = (lambda: torch.cuda.stream(self.comm)) if cuda else contextlib.nullcontext
        cur = torch.cuda.current_stream(self.device) if cuda else None
        e0 = e1 = None
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            self.comm.wait_stream(cur)
        per = max((r_hi - r_lo for r_lo, r_hi in (shard_range(self.n_total, r, self.world) for r in range(self.world))))
        csz = (per + nchunks - 1) // nchunks
        arrived = []
        # scatter pieces (side stream)
        with side():
            for c in range(nchunks):
                ops = []
                for r in range(self.world):
                    a, b = shard_range(self.n_total, r, self.world)
                    pa, pb = min(a + c * csz, b), min(a + (c + 1) * csz, b)
                    if pb <= pa:
                        continue
                    if self.rank == self.src:
                        if r == self.src:
                            self.mine[pa - a:pb - a].copy_(root_frames[pa:pb], non_blocking=True)
                        else:
                            ops.append(dist.P2POp(dist.isend, root_frames[pa:pb], r, self.group))
                    elif r == self.rank:
                        ops.append(dist.P2POp(dist.irecv, self.mine[pa - a:pb - a], self.src, self.group))
                _run(ops)
                if cuda:
                    e = torch.cuda.Event()
                    e.record(self.comm)
                    arrived.append(e)
        # compute pieces (current stream), each followed by its gather on the side stream
        res = self.pipe.results(self.m)
        for c in range(nchunks):
            pa, pb = min(c * csz, self.m), min((c + 1) * csz, self.m)
            if cuda:
                cur.wait_event(arrived[c])
            if pb > pa:
                self.pipe.run(self.mine[pa:pb], pa)
            if cuda:
                e = torch.cuda.Event()
                e.record(cur)
            with side():
                if cuda:
                    self.comm.wait_event(e)
                ops = []
                for r in range(self.world):
                    a, b = shard_range(self.n_total, r, self.world)
                    qa, qb = min(a + c * csz, b), min(a + (c + 1) * csz, b)
                    if qb <= qa:
                        continue
                    for gi, k in enumerate(self.keys):
                        if self.rank == self.src:
                            if r == self.src:
                                self.gathered[gi][qa:qb].copy_(res[k][qa - a:qb - a], non_blocking=True)
                            else:
                                ops.append(dist.P2POp(dist.irecv, self.gathered[gi][qa:qb], r, self.group))
                        elif r == self.rank:
                            ops.append(dist.P2POp(dist.isend, res[k][qa - a:qb - a], self.src, self.group))
                _run(ops)
        if cuda:
            cur.wait_stream(self.comm)
            e1.record(cur)
        return e0, e1
