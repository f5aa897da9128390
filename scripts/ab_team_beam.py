"""Development aid: time the solver call of one device beam-search step (DeviceGEDState.solve over P*W*k*k candidates) under
the three launch hints.   python scripts/ab_team_beam.py [config] [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.device_env import DeviceGEDState
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data

cfg = sys.argv[1] if len(sys.argv) > 1 else "aids-50+"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
env = GEDenv("ipfp", "AIDS-20-30")
pairs = synthetic.make_pairs(cfg, P, seed_offset=501)
print("sizes", [(a.n, b.n) for a, b in pairs])
root = DeviceGEDState(env, [Data.from_host(a, dev) for a, _ in pairs], [Data.from_host(b, dev) for _, b in pairs], device=dev)
cand = root.expand(27)
print("candidates", len(cand), "classes", cand.plan[1] if cand.plan else None)
orig = solver.launch_hint_for
for hint in (1, 2, 0, 2, 1):
    solver.launch_hint_for = lambda n, h=hint: h
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ged = cand.solve()
        e1.record()
        torch.cuda.synchronize()
    print(f"hint {hint}: {e0.elapsed_time(e1):.1f} ms  checksum {float(ged.double().sum())}")
