"""world_size-2 gloo test of the N>1 host path: LPT sharding + token/timing gather (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from t5gemma_tts_b200.sharding import lpt_shard, gather_token_streams, run_sharded


def test_lpt_shard_balances_and_covers():
    rng = np.random.default_rng(0)
    costs = rng.integers(150, 750, 2048).tolist()
    for world in (1, 2, 4, 8):
        sh = lpt_shard(costs, world)
        assert sorted(i for s in sh for i in s) == list(range(2048))
        loads = [sum(costs[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(costs)          # LPT bound
    assert lpt_shard([], 4) == [[], [], [], []]
    assert lpt_shard([5.0], 2) == [[0], []]


def _fake_tokens(i):
    return (np.arange(10 + (i * 7) % 23, dtype=np.int64) * (i + 1)) % 65541


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [10 + (i * 7) % 23 for i in range(37)]
        res, sec, total = run_sharded(lambda idx: [_fake_tokens(i) for i in idx], costs)
        ok = all(np.array_equal(res[i], _fake_tokens(i)) for i in range(37)) and total == sum(costs) and sec >= 0
        # ragged / empty shard edge case: fewer requests than ranks
        res2, _, total2 = run_sharded(lambda idx: [_fake_tokens(i) for i in idx], [3.0])
        ok = ok and np.array_equal(res2[0], _fake_tokens(0)) and total2 == len(_fake_tokens(0))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert out == [(0, True), (1, True)]


def test_gather_single_process():
    res, sec, total = gather_token_streams([1, 0], [np.array([5, 6]), np.array([7])], 0.5, 2)
    assert np.array_equal(res[0], [7]) and np.array_equal(res[1], [5, 6]) and total == 3 and sec == 0.5
