"""Development aid: a tiny team launch (and a tiny one-warp launch) for compute-sanitizer runs.
   compute-sanitizer --tool racecheck python scripts/sanitize_team.py ;  --tool memcheck ; --tool synccheck"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

dev = torch.device("cuda:0")
for cfg, iters in (("aids-30-50", 3), ("n=70", 2)):
    pairs = synthetic.make_pairs(cfg, 3, seed_offset=5)
    edits = synthetic.make_edits(pairs, 2, seed_offset=5)
    ps = PairSet.from_host(pairs, dev)
    a = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], max_iter=iters, check=True, launch_hint=2)
    b = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], max_iter=iters, check=True, launch_hint=0)
    torch.cuda.synchronize()
    assert torch.equal(a.ged, b.ged) and torch.equal(a.rowmatch, b.rowmatch)
    print(cfg, "ok", a.ged.tolist())
