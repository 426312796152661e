"""Image sharding across the GPUs of one box.

The codec path shards by IMAGE: the 16 bands of one image are coupled through halos at every layer, images are
independent.  One process per GPU; image i of a job goes to rank i mod world; weights are replicated.  There is NO
collective on the data path - torch.distributed is used only for the barrier around timed regions and for the
max-over-ranks / sum-over-ranks of scalar measurements (NCCL on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def world_info():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


def shard_indices(n_items, world, rank):
    """Round-robin shard: item i -> rank i mod world.  Shards are disjoint, cover range(n_items) and differ in size by
    at most one item."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, int(n_items), world))


def shard_sizes(n_items, world):
    return [len(range(r, int(n_items), world)) for r in range(world)]


def _reduce(value, op, device):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value, device="cpu"):
    """Slowest rank's time: the job is done when the last shard is done."""
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device="cpu"):
    return _reduce(value, dist.ReduceOp.SUM, device)


def job_throughput(units_this_rank, seconds_this_rank, device="cpu"):
    """Whole-job throughput = units processed by ALL ranks / the slowest rank's time."""
    total = sum_over_ranks(units_this_rank, device)
    slowest = max_over_ranks(seconds_this_rank, device)
    return total / slowest, total, slowest
