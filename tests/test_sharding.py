"""N > 1 host logic on CPU: two gloo ranks shard a batch with no data-path collective; the only
communication is the barrier and the max/sum-over-ranks of scalars that bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optical_flow_1_b200 import shard


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = shard.shard_range(n, r, world)
                assert 0 <= a <= b <= n
                seen.extend(range(a, b))
            assert seen == list(range(n))
            sizes = [shard.shard_range(n, r, world)[1] - shard.shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, npairs, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = shard.shard_range(npairs, rank, world)
    # stand-in for the solver: a per-pair value that depends only on the pair's global seed
    vals = np.array([np.random.RandomState(shard.pair_seed(1234, b)).uniform() for b in range(first, last)])
    shard.barrier()
    fake_ms = 10.0 + 5.0 * rank
    worst = shard.max_over_ranks(fake_ms)
    total = shard.sum_over_ranks(len(vals))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([[first, last, worst, total], vals]))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    world, npairs = 2, 11
    port = _free_port()
    mp.spawn(_worker, args=(world, port, npairs, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    ranges = [(int(g[0]), int(g[1])) for g in got]
    assert ranges == [(0, 6), (6, 11)]
    for g in got:
        assert g[2] == 15.0            # max over ranks of the per-rank time
        assert g[3] == npairs          # units all ranks processed
    merged = np.concatenate([g[4:] for g in got])
    expect = np.array([np.random.RandomState(1234 + b).uniform() for b in range(npairs)])
    assert np.array_equal(merged, expect)   # same pairs, same results, whatever the rank count
