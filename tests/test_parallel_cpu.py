"""N>1 host logic on CPU: two gloo ranks shard clips, run a stand-in per-clip op, and gather the frames
in clip order (the product path does the same with NCCL; bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pgtformer_b200.parallel import gather_frames, shard_clips
    x = torch.arange(n_clips * 3 * 4, dtype=torch.float32).view(n_clips * 3, 4)
    mine = shard_clips(x, rank, world)
    y = mine * 2.0 + 1.0                       # stand-in for the per-clip forward (clips are independent)
    full = gather_frames(y, n_clips)
    ok = torch.equal(full, x * 2.0 + 1.0)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(bool(t.item() == 1.0))
    dist.destroy_process_group()


def _run(n_clips):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return q.get(timeout=10)


def test_shard_ranges():
    from pgtformer_b200.parallel import shard_range
    assert [shard_range(16, r, 8) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert shard_range(1, 1, 2) == (1, 1)


def test_two_rank_gloo_gather_equal_shards():
    assert _run(4)


def test_two_rank_gloo_gather_ragged_shards():
    assert _run(5)
