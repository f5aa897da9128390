"""Development aid: small-batch launch shape A/B (GED_B200_SPREAD)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

dev = torch.device("cuda:0")
for config in ("aids-20-30", "aids-30-50", "aids-50+"):
    for P, E in ((10, 6), (10, 27), (32, 32), (64, 64), (64, 128)):
        pairs = synthetic.make_pairs(config, P)
        edits = synthetic.make_edits(pairs, E)
        ps = PairSet.from_host(pairs, dev)
        out = []
        for spread in (0, 1):
            os.environ["GED_B200_SPREAD"] = str(spread)
            best = 1e9
            for rep in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                res = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2])
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out.append(best)
        print(f"{config} B={len(edits[0]):5d}: packed {out[0]:8.1f} ms  spread {out[1]:8.1f} ms  ({out[0] / out[1]:.2f}x)", flush=True)
