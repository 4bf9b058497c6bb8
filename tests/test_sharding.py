"""World-size-2 gloo test of the data-parallel driver (host logic; runs on CPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_everything():
    from progressivecodec_b200.sharding import shard_bounds

    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from progressivecodec_b200.sharding import run_sharded

    items = list(range(11))
    # stand-in for compress(): returns something that depends on the item and records which rank coded it
    out = run_sharded(lambda i: {"item": i, "rank": rank, "bytes": bytes([i]) * (i + 1)}, items)
    if rank == 0:
        ret.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_run_sharded_world2_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [o["item"] for o in out] == list(range(11))
    assert [o["rank"] for o in out] == [0] * 6 + [1] * 5
    assert all(o["bytes"] == bytes([o["item"]]) * (o["item"] + 1) for o in out)


def test_run_sharded_without_process_group_is_a_map():
    from progressivecodec_b200.sharding import run_sharded

    assert run_sharded(lambda v: v * v, [1, 2, 3]) == [1, 4, 9]
