"""Request sharding for one-engine-replica-per-GPU serving.

Utterances are independent, so the data path has no collective: requests are dealt to ranks longest-first
(LPT) by their token budget, every rank runs continuous batching locally, and torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests) is used only to gather the token streams and timings."""
from __future__ import annotations

import heapq
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def lpt_shard(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of item indices to `world` ranks (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    heap = [(0.0, r) for r in range(world)]
    heapq.heapify(heap)
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + float(costs[i]), r))
    return out


def gather_token_streams(local_idx: Sequence[int], local_tokens: Sequence[np.ndarray], local_seconds: float,
                         n_total: int, device="cpu") -> Tuple[List[np.ndarray], float, int]:
    """All ranks call this.  Returns (token streams in global request order, max seconds over ranks,
    total tokens).  One all_gather of counts, one of padded int32 buffers, one all_reduce(MAX) of time."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        res = [None] * n_total
        for i, t in zip(local_idx, local_tokens):
            res[i] = np.asarray(t)
        return res, float(local_seconds), int(sum(len(t) for t in local_tokens))
    n_local = torch.tensor([len(local_idx)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    max_req = int(max(c.item() for c in counts))
    max_len = torch.tensor([max((len(t) for t in local_tokens), default=0)], dtype=torch.int64, device=device)
    dist.all_reduce(max_len, op=dist.ReduceOp.MAX)
    L = int(max_len.item())
    buf = torch.full((max_req, L + 2), -1, dtype=torch.int32, device=device)       # [idx, len, tokens...]
    for j, (i, t) in enumerate(zip(local_idx, local_tokens)):
        buf[j, 0] = i
        buf[j, 1] = len(t)
        if len(t):
            buf[j, 2: 2 + len(t)] = torch.as_tensor(np.asarray(t, dtype=np.int32), device=device)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    sec = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    res: List[np.ndarray] = [None] * n_total        # type: ignore
    total = 0
    for r in range(world):
        b = bufs[r].cpu().numpy()
        for j in range(int(counts[r].item())):
            i, n = int(b[j, 0]), int(b[j, 1])
            res[i] = b[j, 2: 2 + n].astype(np.int64)
            total += n
    return res, float(sec.item()), total


def run_sharded(generate: Callable[[List[int]], List[np.ndarray]], costs: Sequence[float], device="cpu"):
    """generate(indices) -> token arrays for those request indices (this rank's engine).  Returns the gathered
    result on every rank."""
    import time
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = lpt_shard(costs, world)[rank]
    if dist.is_initialized():
        dist.barrier()
    t0 = time.perf_counter()
    toks = generate(mine)
    if torch.cuda.is_available() and str(device) != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return gather_token_streams(mine, toks, dt, len(costs), device=device)
