"""world_size-2 gloo test of the N>1 path's host logic (sharding + the single all-gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensorflow_web_deploy_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 2, 5, 128, 129, 1024, 1027):
        for parts in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, parts, r) for r in range(parts)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sharding.shard_bounds(n, parts, parts) == (n, n)


def test_gpus_for_batch_policy():
    assert sharding.gpus_for_batch(1, 8) == 1           # configs[1]: a single request stays on one GPU
    assert sharding.gpus_for_batch(128, 8) == 1
    assert sharding.gpus_for_batch(129, 8) == 2
    assert sharding.gpus_for_batch(1024, 8) == 8        # configs[3]: 128 per GPU
    assert sharding.gpus_for_batch(4096, 8) == 8
    assert sharding.gpus_for_batch(1024, 2) == 2


def _fake_forward(x):
    # deterministic per-row "probabilities": depends on the row content only
    s = x.reshape(x.shape[0], -1).sum(1, dtype=np.float64)
    return np.stack([np.sin(s), np.cos(s), s % 7.0], 1).astype(np.float32)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        rng = np.random.default_rng(0)
        for n, sweet in ((5, 2), (4, 2), (1, 2), (2, 128), (7, 1), (3, 1)):
            x = rng.standard_normal((n, 4, 4, 3)).astype(np.float32)
            calls = []
            tower = sharding.ShardedTower(lambda r: (calls.append(r.shape[0]), _fake_forward(r))[1], 3, dist=dist,
                                          sweet_spot=sweet)
            out = tower.classify(x)
            res[(n, sweet)] = (out, calls, _fake_forward(x))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_rank_all_gather_reassembles_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        for (n, sweet), (out, calls, ref) in got[rank].items():
            np.testing.assert_array_equal(out, ref)                     # every rank holds the full, ordered result
            parts = sharding.gpus_for_batch(n, 2, sweet)
            lo, hi = sharding.shard_bounds(n, parts, rank)
            assert calls == ([hi - lo] if hi > lo else [])               # each rank computed only its rows
    assert got[1][(2, 128)][1] == []                                     # below the sweet spot: rank 1 idles
