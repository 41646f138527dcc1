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


def restored_middle_u8(out):
    """What the consumer of the path keeps (`inference.py:15-19`): the middle frame of every clip as
    uint8(clamp(x, 0, 1) * 255), rgb24 [clips, H, W, 3] — computed on the device (12x fewer bytes than fp32 `out`)."""
    from . import ops
    n = out.shape[0] // 3
    return ops.f32nchw_to_u8hwc(out, torch.empty(n, out.shape[2], out.shape[3], 3, dtype=torch.uint8, device=out.device),
                                first=1, step=3)


def gather_restored(out, n_clips):
    """The path's one collective in the form the consumer needs: all-gather of the restored middle frames (rgb24)."""
    return gather_frames(restored_middle_u8(out), n_clips, frames_per_clip=1)
