"""N>1 host logic on CPU: world_size-2 gloo processes exercise the stream partition, the max-over-ranks
timing reduction and the whole-job throughput formula bench.py uses under torchrun."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zaru_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.streams_for_rank(n_streams, world, rank)
    seconds = 0.5 + 0.25 * rank                      # rank 1 is the slow one
    tput = shard.aggregate_throughput(len(mine) * 10.0, seconds, dist)
    (slowest,) = shard.max_over_ranks([seconds], dist)
    everyone = shard.gather_records({"rank": rank, "streams": mine}, dist)
    dist.barrier()
    q.put((rank, mine, tput, slowest, everyone))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [64, 7])
def test_world2_partition_and_reduction(n_streams):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_streams = sorted(s for _, mine, *_ in results for s in mine)
    assert all_streams == list(range(n_streams))                       # disjoint cover
    for rank, mine, tput, slowest, everyone in results:
        assert all(s % world == rank for s in mine)
        assert slowest == pytest.approx(0.75)                          # max over ranks, not this rank's own time
        assert tput == pytest.approx(n_streams * 10.0 / 0.75)          # all ranks' units / slowest rank
        assert [e["rank"] for e in everyone] == [0, 1]
        assert sorted(s for e in everyone for s in e["streams"]) == list(range(n_streams))


def test_single_process_degenerates():
    assert shard.streams_for_rank(5, 1, 0) == [0, 1, 2, 3, 4]
    assert shard.aggregate_throughput(100.0, 2.0) == 50.0
    assert shard.gather_records([1, 2]) == [[1, 2]]
    with pytest.raises(ValueError):
        shard.streams_for_rank(4, 2, 2)
