"""CPU, world_size 2 over gloo: the N>1 path of the candidate sharding (partition -> per-rank results -> gather in
candidate order -> identical to the single-process result; small-tensor mean all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ged_oracle as go
from ppo_bihyb_b200 import sharding, synthetic
from ppo_bihyb_b200.graphs import label_node_cost


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _candidates():
    pairs = synthetic.make_pairs("aids-20-30", 3, seed_offset=21)
    base, eu, ev = synthetic.make_edits(pairs, 5, seed_offset=4)
    return pairs, base, eu, ev


def _solve_subset(pairs, base, eu, ev, idx):
    """Stand-in for the per-rank GPU solve: the CPU oracle (hungarian start only, to keep the test fast)."""
    out = []
    for b in idx:
        g1, g2 = pairs[base[b]]
        o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), edit=(int(eu[b]), int(ev[b])), max_iter=3)
        out.append(o["ged"])
    return torch.tensor(out, dtype=torch.float32)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs, base, eu, ev = _candidates()
        B = len(base)
        n1 = np.asarray([pairs[p][0].n for p in base])
        n2 = np.asarray([pairs[p][1].n for p in base])
        for cost in (None, sharding.work_estimate(n1, n2)):
            idx = sharding.shard_indices(B, rank, world, cost)
            local = _solve_subset(pairs, base, eu, ev, idx)
            full = sharding.gather_results(local, idx, B)
            q.put((rank, full.numpy().copy()))
        grads = [torch.full((3, 2), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))]
        sharding.allreduce_mean_(grads)
        q.put((rank, np.concatenate([g.reshape(-1).numpy() for g in grads])))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather_equals_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world * 3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pairs, base, eu, ev = _candidates()
    want = _solve_subset(pairs, base, eu, ev, range(len(base))).numpy()
    fulls = [v for _, v in got if v.shape == want.shape]
    assert len(fulls) == 4
    for v in fulls:
        assert np.array_equal(v, want)                    # bitwise: sharding must not change any result
    means = [v for _, v in got if v.shape != want.shape]
    for v in means:
        assert np.allclose(v[:6], 1.5) and np.allclose(v[6:], 15.0)
