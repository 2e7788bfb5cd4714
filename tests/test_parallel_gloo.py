"""Multi-process (world_size 2, gloo, CPU) tests of the host-side sharding logic: envs shard as contiguous
slabs with global env ids, and the only cross-rank data is the episode-statistics vector."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from madigan_b200 import _abi as A
from madigan_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_envs_partition():
    for total in (1, 7, 65_536, 1_048_576, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_envs(total, world, r) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == total
            for (o0, c0), (o1, _c1) in zip(spans, spans[1:]):
                assert o0 + c0 == o1  # contiguous, ordered
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, n_assets, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # what each rank's mdg_episode_stats would produce for its slab (synthetic, deterministic)
    off, cnt = parallel.shard_envs(1000, world, rank)
    eq = 1e6 + np.arange(off, off + cnt, dtype=np.float64)
    v = torch.zeros(A.MDG_STATS_NSCALAR + 2 * n_assets, dtype=torch.float64)
    v[0], v[1], v[2], v[3], v[4] = cnt, eq.sum(), (eq ** 2).sum(), eq.min(), eq.max()
    v[5], v[6], v[7] = 0.001 * cnt, 2.0 * cnt, rank
    v[A.MDG_STATS_NSCALAR:] = rank + 1
    red = parallel.reduce_episode_stats(v, n_assets)
    q.put((rank, red.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_episode_stats_world2():
    world, n_assets = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_assets, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eq = 1e6 + np.arange(1000, dtype=np.float64)
    for r in range(world):
        v = res[r]
        assert v[0] == 1000
        np.testing.assert_allclose(v[1], eq.sum(), rtol=1e-15)
        np.testing.assert_allclose(v[2], (eq ** 2).sum(), rtol=1e-15)
        assert v[3] == eq.min() and v[4] == eq.max()
        assert v[7] == 0 + 1
        np.testing.assert_allclose(v[A.MDG_STATS_NSCALAR:], 3.0)
    s = parallel.summarize_stats(torch.from_numpy(res[0]), n_assets)
    assert s["n_envs"] == 1000 and abs(s["mean_equity"] - eq.mean()) < 1e-6
    assert abs(s["std_equity"] - eq.std()) < 1e-3
