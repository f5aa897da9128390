"""Short single-launch driver for ncu captures of ged_solve_kernel (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

config = sys.argv[1] if len(sys.argv) > 1 else "n=60"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 32
E = int(sys.argv[3]) if len(sys.argv) > 3 else 32
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
pairs = synthetic.make_pairs(config, P)
edits = synthetic.make_edits(pairs, E)
ps = PairSet.from_host(pairs, dev)
for rep in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    B = len(edits[0])
    print(f"{config}: B={B} {ms:.2f} ms -> {B / ms * 1e3:.0f} solves/s; iters {res.iters.float().mean().item():.1f}; "
          f"bad {(res.status != 0).sum().item()}; ged {res.ged.mean().item():.3f}", flush=True)
