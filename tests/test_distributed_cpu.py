"""N > 1 host logic on CPU: two gloo ranks shard the env index space and reduce the episode
statistics vector (the only collective the framework issues)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from competitive_rl_b200.distributed import gather_episode_stats, shard_range, stats_from_raw
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(total, rank, world)
    # fake per-shard statistics: one finished episode per owned env, length = global env index
    idx = list(range(first, first + count))
    local = stats_from_raw([len(idx), sum(idx), sum(1 for i in idx if i % 3 == 0), sum(1 for i in idx if i % 3 == 1),
                            sum(1 for i in idx if i % 3 == 2), 64 * len(idx) + sum(i % 5 - 2 for i in idx), 0, 0])
    whole = gather_episode_stats(local)
    out[rank] = (first, count, whole)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 1001), (2, 64)])
def test_shard_and_gather_gloo(world, total):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    firsts = sorted((out[r][0], out[r][1]) for r in range(world))
    # shards tile [0, total) without gaps or overlap
    pos = 0
    for f, c in firsts:
        assert f == pos
        pos += c
    assert pos == total
    idx = range(total)
    for r in range(world):
        w = out[r][2]
        assert w["episodes"] == total and w["sum_length"] == sum(idx)
        assert w["left_wins"] + w["right_wins"] + w["draws"] == total
        assert abs(w["mean_margin"] - sum(i % 5 - 2 for i in idx) / total) < 1e-12


def test_shard_range_edges():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from competitive_rl_b200.distributed import shard_range
    assert shard_range(8, 0, 8) == (0, 1) and shard_range(8, 7, 8) == (7, 1)
    assert shard_range(3, 2, 4) == (2, 1) and shard_range(3, 3, 4) == (3, 0)
    assert shard_range(65536 * 8, 5, 8) == (5 * 65536, 65536)
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)
