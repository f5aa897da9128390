"""The benchmarked workload of bench.py (same pairs, same candidate list, same plan), `reps` steps and nothing else: the
command ncu is pointed at for the per-class instruction counts of the roofline (profiles/roofline_r02.json).
    python scripts/prof_bench_step.py [reps] [--config aids-50+ --pairs 64 --edits 256]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ppo_bihyb_b200 import solver
from ppo_bihyb_b200.graphs import PairSet

ap = argparse.ArgumentParser()
ap.add_argument("reps", nargs="?", type=int, default=2)
ap.add_argument("--config", default="aids-50+")
ap.add_argument("--pairs", type=int, default=64)
ap.add_argument("--edits", type=int, default=256)
ap.add_argument("--max-iter", type=int, default=100)
ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
ap.add_argument("--total", type=int, default=131072)
args = ap.parse_args()
dev = torch.device("cuda:0")
pairs, base, eu, ev = bench.workload(args, 0)
ps = PairSet.from_host(pairs, dev)
plan = solver.make_plan(ps, base)
b_d, u_d, v_d = (torch.from_numpy(a).to(dev) for a in (base, eu, ev))
for rep in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = solver.solve_batch(ps, base_idx=b_d, edit_u=u_d, edit_v=v_d, max_iter=args.max_iter, plan=plan, lap_steps=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {rep}: {e0.elapsed_time(e1):.1f} ms, {len(base)} candidates in {res.launches} launches {plan[1]}, "
          f"{int(res.trace['lap_steps'].sum())} LAP scan steps, failed {int((res.status != 0).sum())}", flush=True)
