"""World-size-2 gloo tests (CPU) of the N>1 path: env sharding and the episode-statistics all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from isaac_rover_orbit_b200 import dist as RD


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = RD.shard_range(n_total, rank, world)
        g = torch.Generator().manual_seed(100)
        # every rank derives the same global per-env data and keeps its shard, like replicated terrain + sharded envs
        sums = torch.rand(n_total, 7, generator=g)
        reset = torch.rand(n_total, generator=g) < 0.3
        flags = (torch.rand(n_total, 4, generator=g) < 0.25) & reset[:, None]
        err = torch.rand(n_total, 2, generator=g)
        local = torch.zeros(RD.STATS_LEN)
        m = reset[lo:hi]
        local[:7] = sums[lo:hi][m].sum(0)
        local[7:11] = flags[lo:hi].float().sum(0)
        local[11:13] = err[lo:hi][m].sum(0)
        local[13] = m.sum()
        stats = RD.EpisodeStats(local.clone(), world)
        log = stats.all_reduce_async().log()
        # the accumulator is cleared for the next interval
        assert float(stats.stats.abs().sum()) == 0.0
        if rank == 0:
            expect = RD.episode_log(torch.cat([sums[reset].sum(0), flags.float().sum(0), err[reset].sum(0),
                                               reset.sum()[None].float(), torch.zeros(2)]))
            for k, v in expect.items():
                assert abs(log[k] - v) <= 1e-5 * max(1.0, abs(v)), (k, log[k], v)
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_the_env_axis():
    for n, w in ((16384, 8), (10, 3), (7, 8), (8192 * 8, 8)):
        spans = [RD.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert RD.spawn_rows(8 * 8192, 3, 8) == (2 * 8192 * 3, 2 * 8192 * 4)
    # uneven shards: the row ranges follow shard_range -- disjoint, gap-free, 2 rows per env
    n, world = 1003, 8
    rows = [RD.spawn_rows(n, r, world) for r in range(world)]
    assert rows[0][0] == 0 and rows[-1][1] == 2 * n
    for r in range(world):
        lo, hi = RD.shard_range(n, r, world)
        assert rows[r] == (2 * lo, 2 * hi) and (r == 0 or rows[r - 1][1] == rows[r][0])


def test_episode_stats_sum_over_count_not_mean_of_means():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"


def test_episode_log_single_rank():
    s = torch.zeros(RD.STATS_LEN)
    s[0], s[7], s[11], s[13] = 3.0, 2.0, 8.0, 4.0
    log = RD.episode_log(s)
    assert abs(log["Episode Reward/distance_to_target"] - 3.0 / 4 / 150.0) < 1e-9
    assert log["Episode Termination/time_limit"] == 2 and abs(log["Metrics/target_pose/error_pos"] - 2.0) < 1e-9
    stats = RD.EpisodeStats(s.clone(), 1)
    assert torch.equal(stats.all_reduce_async().result(), s)
