"""Candidate sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Every pair-solve is independent (reference ``GEDenv.step`` is a pure function of its arguments,
``utils/ged_env.py:110-124``), so the candidate list is partitioned with NO data-path collective: each rank solves
its own contiguous block and only the B x 4-byte objectives (optionally the assignments) are gathered afterwards.
The reference's own parallelism is a fork pool mapping ``ged_env.step`` over the batch
(``ged_ppo_bihyb_train.py:259-260,295``); this replaces it.
"""
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def work_estimate(n1: np.ndarray, n2: np.ndarray) -> np.ndarray:
    """Relative cost of a solve: LAP scan steps grow ~ (n1+n2)^2 and each step touches ~ (n1+n2) columns."""
    n = (np.asarray(n1, np.float64) + np.asarray(n2, np.float64))
    return n ** 3


def balanced_blocks(cost: np.ndarray, world_size: int) -> np.ndarray:
    """Assign each candidate to a rank so that the summed work estimate is balanced: candidates sorted by descending
    cost are dealt to the currently lightest rank (LPT).  Returns ``owner`` int32 [B]; deterministic (stable sort, ties
    to the lowest rank)."""
    cost = np.asarray(cost, np.float64)
    B = cost.shape[0]
    owner = np.zeros(B, np.int32)
    if world_size <= 1 or B == 0:
        return owner
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world_size, np.float64)
    count = np.zeros(world_size, np.int64)
    cap = -(-B // world_size)                       # keep counts within one of each other: fixed-size gather buffers
    for b in order:
        open_ranks = np.nonzero(count < cap)[0]
        r = int(open_ranks[np.argmin(load[open_ranks])])
        owner[b] = r
        load[r] += cost[b]
        count[r] += 1
    return owner


def shard_indices(B: int, rank: int, world_size: int, cost: Optional[np.ndarray] = None,
                  interleave: bool = False) -> np.ndarray:
    """Indices (ascending) of the candidates rank ``rank`` solves.  Without a cost estimate: contiguous blocks of
    ceil(B / world_size); with one: the LPT partition of :func:`balanced_blocks`; ``interleave``: candidate b goes to
    rank b % world_size -- the right choice when the list is grouped by base pair (a beam-search fan-out: the edits of one
    pair cost almost the same, so every rank gets the same mix of pairs whatever the error of a size-based estimate)."""
    if interleave:
        return np.arange(rank, B, world_size, dtype=np.int64)
    if cost is None:
        per = -(-B // world_size) if B else 0
        return np.arange(min(B, rank * per), min(B, (rank + 1) * per), dtype=np.int64)
    return np.nonzero(balanced_blocks(cost, world_size) == rank)[0].astype(np.int64)


def gather_results(local: torch.Tensor, local_idx: np.ndarray, B: int, group=None) -> torch.Tensor:
    """All-gather per-rank results back into candidate order.  ``local`` is [len(local_idx), ...] on this rank's device;
    returns [B, ...] on every rank.  Uses ``all_gather_into_tensor`` on fixed-size (padded) blocks: B x 4 bytes of
    objectives is latency-bound on NVLink, so one padded collective beats variable-size ones."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        out[torch.as_tensor(local_idx, device=local.device)] = local
        return out
    world = dist.get_world_size(group)
    cap = -(-B // world)
    tail = tuple(local.shape[1:])
    pad = torch.zeros((cap,) + tail, dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    idx = torch.full((cap,), -1, dtype=torch.int64, device=local.device)
    idx[:len(local_idx)] = torch.as_tensor(local_idx, dtype=torch.int64, device=local.device)
    all_val = torch.empty((world * cap,) + tail, dtype=local.dtype, device=local.device)
    all_idx = torch.empty(world * cap, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(all_val, pad, group=group)
    dist.all_gather_into_tensor(all_idx, idx, group=group)
    keep = all_idx >= 0
    out = torch.empty((B,) + tail, dtype=local.dtype, device=local.device)
    out[all_idx[keep]] = all_val[keep]
    return out


def allreduce_mean_(tensors: List[torch.Tensor], group=None) -> None:
    """In-place mean all-reduce of a list of small tensors (the policy's 290k fp32 gradients, 1.16 MB) as ONE flattened
    NCCL call -- latency-bound, so bucketing is by launch count, not link bandwidth."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    o = 0
    for t in tensors:
        t.copy_(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
