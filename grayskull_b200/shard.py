"""Frame-batch sharding across the GPUs of one box.

Frames are independent units (every gs_* op takes one image), so the only multi-GPU data
movement the path has is one scatter of uint8 frames from the root rank and one gather of the
results -- grouped send/recv over NCCL (NVLink 5 / NVSwitch) in production, gloo in the CPU
tests.  One process per GPU; nothing else crosses ranks.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous frame range [lo, hi) of `rank`: GPU g gets frames [g*n/G, (g+1)*n/G)"""
    return (n * rank) // world, (n * (rank + 1)) // world


def scatter_frames(frames, n, frame_shape, dtype, device, src=0, group=None):
    """Root holds `frames` (n, *frame_shape); every rank returns its contiguous shard.
    Implemented as grouped point-to-point sends (NCCL has no native scatter for ragged shards)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n, rank, world)
    mine = torch.empty((hi - lo,) + tuple(frame_shape), dtype=dtype, device=device)
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
                ops.append(dist.P2POp(dist.isend, frames[a:b].contiguous(), r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, mine, src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return mine


def gather_frames(mine, n, dst=0, group=None):
    """Inverse of scatter_frames: root returns (n, ...) assembled in frame order, others None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return mine
    out = None
    ops = []
    if rank == dst:
        out = torch.empty((n,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        for r in range(world):
            a, b = shard_range(n, r, world)
            if r == dst:
                out[a:b].copy_(mine)
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, out[a:b], r, group))
    elif mine.shape[0] > 0:
        ops.append(dist.P2POp(dist.isend, mine.contiguous(), dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out
