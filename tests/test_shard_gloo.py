"""CPU suite: the N>1 host logic — contiguous frame sharding and the final all-gather of outputs — on
world_size = 2 (and a ragged 3-way split) with the gloo backend.  No GPU, no kernels: this covers what
`bench.py --gpus N` and `SmirkPipeline` do around the per-rank hot path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smirk_b200.pipeline import all_gather_frames, shard_bounds


def test_shard_bounds_cover_the_batch_exactly():
    for n, ws in ((64, 2), (2048, 8), (10, 3), (1, 4), (0, 2), (7, 7)):
        spans = [shard_bounds(n, ws, r) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    assert shard_bounds(64, 2, 1) == (32, 64)            # configs[3]: 64 frames over 2 GPUs
    assert shard_bounds(2048, 8, 3) == (768, 1024)       # configs[4]: 256 per GPU


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_frames * 3 * 4, dtype=torch.float32).view(n_frames, 3, 4)      # "frames"
        lo, hi = shard_bounds(n_frames, world, rank)
        local = full[lo:hi] * 2.0 + 1.0                                                      # per-rank "hot path"
        got = all_gather_frames(local, n_frames)
        ok = torch.equal(got, full * 2.0 + 1.0)
        # barrier + max-over-ranks timing reduction, as bench.py does it
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bool(ok), float(t.item()), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 64), (2, 7), (3, 10)])
def test_all_gather_frames_gloo(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, tmax, shape in res:
        assert ok, "rank %d gathered a wrong batch" % rank
        assert tmax == float(world) and shape == (n_frames, 3, 4)
