"""Row f-4 measurement: batched RRWM / GA baselines (time per call and per pair)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import baselines, solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

dev = torch.device("cuda:0")
print("| config | pairs | rrwm_batch s | ms / pair | mean GED (rrwm) | ga_batch s | mean GED (ga) | mean GED (ipfp) |")
print("|---|---|---|---|---|---|---|---|")
for config, P in (("aids-20-30", 50), ("aids-30-50", 50), ("aids-50+", 50)):
    pairs = synthetic.make_pairs(config, P, seed_offset=321)
    ps = PairSet.from_host(pairs, dev)
    out = []
    for fn in (baselines.rrwm_batch, baselines.ga_batch):
        fn(ps, max_iter=2)
        torch.cuda.synchronize()
        t0 = time.time()
        r = fn(ps)
        ged = solver.score_batch(ps, r[0], r[1])
        torch.cuda.synchronize()
        out.append((time.time() - t0, float(ged.mean())))
    ipfp = float(solver.solve_batch(ps).ged.mean())
    print(f"| {config} | {P} | {out[0][0]:.2f} | {out[0][0] / P * 1e3:.1f} | {out[0][1]:.1f} | {out[1][0]:.2f} | {out[1][1]:.1f} | {ipfp:.1f} |", flush=True)
