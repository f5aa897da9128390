"""Soak test of the CTA-per-candidate launches (development aid): random candidates per size class solved with launch_hint 2
(team), several times under different load (one wave, persistent CTAs, classes side by side) against the one-warp packed
launch, bit for bit -- objective on both graphs, assignment, iteration count, LAP scan steps.  Usage: soak_team.py [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
bad_total = 0
for config, P, E in (("aids-30-50", 24, 8), ("aids-30-50", 12, 160), ("aids-50+", 48, 6), ("aids-50+", 16, 100), ("n=70", 8, 30),
                     ("n=85", 8, 40), ("n=100", 6, 60), ("n=130", 3, 20), ("n=33", 6, 40), ("n=64", 6, 60), ("n=65", 6, 30), ("n=96", 4, 30),
                     ("n=97", 4, 20), ("n=128", 2, 10), ("n=129", 2, 6)):
    pairs = synthetic.make_pairs(config, P, seed_offset=777)
    edits = synthetic.make_edits(pairs, E, seed_offset=777)
    ps = PairSet.from_host(pairs, dev)
    kw = dict(base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], check=True, lap_steps=True)
    ref = solver.solve_batch(ps, launch_hint=0, **kw)
    bad = 0
    t0 = time.time()
    for rep in range(reps):
        got = solver.solve_batch(ps, launch_hint=2, **kw)
        same = (torch.equal(got.ged, ref.ged) and torch.equal(got.ged_edited, ref.ged_edited) and torch.equal(got.rowmatch, ref.rowmatch)
                and torch.equal(got.colmatch, ref.colmatch) and torch.equal(got.iters, ref.iters)
                and torch.equal(got.trace["lap_steps"], ref.trace["lap_steps"]))
        bad += 0 if same else int((got.ged != ref.ged).sum().item()) + 1
    torch.cuda.synchronize()
    bad_total += bad
    print(f"{config}: {len(edits[0])} candidates x {reps} team runs, {bad} mismatches, {time.time() - t0:.2f} s", flush=True)
print("SOAK TEAM", "FAILED" if bad_total else "OK")
sys.exit(1 if bad_total else 0)
