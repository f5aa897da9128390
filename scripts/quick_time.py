"""Ad-hoc device timing of ged_solve_batch (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

dev = torch.device("cuda:0")
for config, P, E in (("aids-20-30", 64, 64), ("aids-30-50", 64, 64), ("aids-50+", 64, 64)):
    pairs = synthetic.make_pairs(config, P)
    edits = synthetic.make_edits(pairs, E)
    ps = PairSet.from_host(pairs, dev)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    B = len(edits[0])
    print(f"{config}: B={B} {ms:.2f} ms -> {B / ms * 1e3:.0f} solves/s; mean iters {res.iters.float().mean().item():.1f}; "
          f"status nonzero {(res.status != 0).sum().item()}; mean ged {res.ged.mean().item():.2f}", flush=True)
