"""Partitioning of independent pictures / streams over the GPUs of one box.

Intra pictures and separate streams share nothing, so the path shards by picture or
stream with no data-path collective: stream s -> rank s mod world.  The
only inter-rank traffic is the benchmark's barrier and the MAX reduction of the timed
region: host scalars over torch.distributed's gloo backend -- no NCCL anywhere.  Inside one
process the same partition is what `pool.EnginePool` dispatches by."""
from __future__ import annotations


def streams_of_rank(n_streams: int, rank: int, world: int):
    """Stream ids handled by `rank` (round robin)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    return list(range(rank, n_streams, world))


def pictures_of_rank(n_pics: int, rank: int, world: int):
    """Picture p -> rank p mod world (used when one stream is split by picture)."""
    return streams_of_rank(n_pics, rank, world)


def max_over_ranks(value: float, device=None) -> float:
    """MAX of a host scalar over all ranks (identity without a process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_rate(units_this_rank: float, seconds_this_rank: float, device=None) -> float:
    """Aggregate throughput of the job: units of all ranks / slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
