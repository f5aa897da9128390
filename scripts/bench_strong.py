"""Strong scaling of the candidate-edit solve: a FIXED total candidate list -- 64 AIDS-50+
base pairs x (total / 64) edge toggles -- sharded over the ranks with no data-path collective; the B x 4-byte objectives are
all-gathered at the end of a step.  Device-timed (CUDA events), max over ranks, L2 flushed between steps.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_strong.py [--totals 131072,32768,4096]

Prints one JSON line per total (rank 0)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from ppo_bihyb_b200 import sharding, solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

ap = argparse.ArgumentParser()
ap.add_argument("--totals", default="131072,32768,4096")
ap.add_argument("--config", default="aids-50+")
ap.add_argument("--pairs", type=int, default=64)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pairs = synthetic.make_pairs(args.config, args.pairs, seed_offset=100)
ps = PairSet.from_host(pairs, dev)
for total in (int(t) for t in args.totals.split(",")):
    base, eu, ev = synthetic.make_edits(pairs, total // args.pairs, seed_offset=100)
    idx = sharding.shard_indices(len(base), rank, world, interleave=True) if world > 1 else np.arange(len(base))
    b_d, u_d, v_d = (torch.from_numpy(a[idx]).to(dev) for a in (base, eu, ev))
    plan = solver.make_plan(ps, base[idx])
    B, B_max = len(idx), -(-len(base) // world)

    def step():
        res = solver.solve_batch(ps, base_idx=b_d, edit_u=u_d, edit_v=v_d, plan=plan)
        if world > 1:
            mine = torch.zeros(B_max, dtype=torch.float32, device=dev)
            mine[:B] = res.ged
            allged = torch.empty(world * B_max, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allged, mine)
        return res

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = step()
        e1.record()
        e1.synchronize()
        ms += e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    bad = torch.tensor([int((res.status != 0).sum())], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(bad)
    if rank == 0:
        print(json.dumps({"metric": "ipfp_ged_pair_solves_per_sec", "scaling": "strong", "n_gpus": world, "total_candidates": len(base),
                          "candidates_per_gpu": B_max, "value": len(base) * args.steps / (float(t[0]) * 1e-3),
                          "ms_per_step": float(t[0]) / args.steps, "failed_solves": int(bad[0]), "config": args.config}), flush=True)
if world > 1:
    dist.destroy_process_group()
