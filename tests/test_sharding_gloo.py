"""Multi-process host logic of the image-sharded path, world_size 2 over gloo on CPU (the GPU path runs the same code
over NCCL): shards partition the job, the whole-job throughput is total units / slowest rank, and the in-memory coder
gives rank-independent bitstreams."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pseudocylindrical_convolution_b200 import sharding


def test_shards_partition_the_job():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            shards = [sharding.shard_indices(n, world, r) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(n))
            sizes = [len(s) for s in shards]
            assert sizes == sharding.shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.world_info() == (rank, rank, world)
        mine = sharding.shard_indices(5, world, rank)                    # 5 images over 2 ranks: 3 + 2
        seconds = 1.0 + rank                                             # rank 1 is the slow one
        dist.barrier()
        thr, total, slowest = sharding.job_throughput(len(mine), seconds)
        # per-image bitstreams must not depend on which rank coded them
        from pseudocylindrical_convolution_b200 import coder
        digests = {}
        for i in mine:
            rng = np.random.default_rng(100 + i)
            n = 200
            cuts = np.sort(rng.integers(1, 65535, size=(n, 7)), axis=1)
            cuts += np.arange(7)                                         # strictly increasing interior boundaries
            table = np.concatenate([np.zeros((n, 1), np.int64), cuts, np.full((n, 1), 65536 + 8)], axis=1).astype(np.int32)
            sym = rng.integers(0, 8, size=n).astype(np.int32)
            c = coder.coder("/tmp/pcx_gloo_%d_%d.bin" % (os.getpid(), i))
            c.start_encoder()
            c.encodes(torch.from_numpy(table), 8, torch.from_numpy(sym), n)
            c.end_encoder()
            digests[i] = open("/tmp/pcx_gloo_%d_%d.bin" % (os.getpid(), i), "rb").read()
            os.remove("/tmp/pcx_gloo_%d_%d.bin" % (os.getpid(), i))
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, thr, total, slowest, {k: len(v) for k, v in digests.items()}))
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered = out.get()
    assert sorted(i for g in gathered for i in g[0]) == [0, 1, 2, 3, 4]
    for mine, thr, total, slowest, sizes in gathered:
        assert total == 5 and slowest == 2.0 and abs(thr - 2.5) < 1e-12      # same answer on every rank
        assert all(v > 16 for v in sizes.values())
