"""N>1 host logic on CPU: world_size-2 gloo processes shard frames, agree on max/sum."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparse_pooling_b200 import sharding


def test_frames_for_rank_partitions_exactly():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in sharding.frames_for_rank(n, r, world)]
            assert got == list(range(n))
            sizes = [len(sharding.frames_for_rank(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frames_for_rank(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = list(sharding.frames_for_rank(5, rank, world))
        t_max = sharding.max_over_ranks(10.0 + rank)
        n_sum = sharding.sum_over_ranks(len(mine))
        recs = sharding.gather_records({"rank": rank, "frames": mine})
        q.put((rank, mine, t_max, n_sum, recs))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == [0, 1, 2] and results[1][1] == [3, 4]
    for r in results:
        assert r[2] == 11.0 and r[3] == 5.0
        assert [x["frames"] for x in r[4]] == [[0, 1, 2], [3, 4]]
