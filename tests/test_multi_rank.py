"""N > 1 host logic on CPU: two ranks over gloo -- partition of streams, barrier, MAX of
the timed region, whole-job aggregate (what bench.py does over NCCL on the GPU box)."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from p265_b200 import partition


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        streams = partition.streams_of_rank(8, rank, world)
        dist.barrier()
        seconds = 1.0 + rank                       # rank 1 is the slow one
        units = 100.0 * len(streams)
        rate = partition.whole_job_rate(units, seconds)
        out.put((rank, streams, partition.max_over_ranks(seconds), rate))
    finally:
        dist.destroy_process_group()


def test_two_ranks_partition_and_reduce():
    ctx = mp.get_context("spawn")
    out, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    assert sorted(res[0][1] + res[1][1]) == list(range(8))          # every stream exactly once
    assert res[0][2] == res[1][2] == 2.0                           # max over ranks
    assert res[0][3] == res[1][3] == pytest.approx(800.0 / 2.0)    # whole job / slowest rank


def test_partition_edge_cases():
    assert partition.streams_of_rank(8, 0, 1) == list(range(8))
    assert partition.streams_of_rank(3, 3, 8) == []
    assert partition.pictures_of_rank(5, 1, 2) == [1, 3]
    with pytest.raises(ValueError):
        partition.streams_of_rank(8, 2, 2)
    assert partition.max_over_ranks(3.5) == 3.5                    # no process group: identity
