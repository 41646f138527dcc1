"""Data-parallel plumbing for the clip-sharded forward: clips are independent, so the only
collective on the path is the final gather of the restored frames.  One process per GPU
(`torch.distributed`, NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests)."""
import torch
import torch.distributed as dist


def shard_range(n_clips, rank, world):
    """Contiguous block of clips owned by `rank` (blocks differ by at most one clip)."""
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_clips(x, rank, world, frames_per_clip=3):
    """x: [n_clips*3, ...] -> this rank's frames (a view)."""
    n_clips = x.shape[0] // frames_per_clip
    s, e = shard_range(n_clips, rank, world)
    return x[s * frames_per_clip:e * frames_per_clip]


def gather_frames(local, n_clips, frames_per_clip=3):
    """All-gathers per-rank outputs [local_clips*3, ...] into [n_clips*3, ...] in clip order.
    Equal shards use one all_gather_into_tensor; ragged shards are padded to the largest shard."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_clips, r, world) for r in range(world)]
    counts = [(e - s) * frames_per_clip for s, e in sizes]
    mx = max(counts)
    if min(counts) == mx:
        out = local.new_empty((mx * world,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    buf = local.new_empty((mx * world,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, pad)
    return torch.cat([buf[r * mx:r * mx + counts[r]] for r in range(world)], 0)
