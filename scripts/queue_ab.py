"""Development aid: A/B of the persistent work queue (GED_B200_WORK_QUEUE) and of the hybrid CTA's shared-memory warps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

dev = torch.device("cuda:0")
ks = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [-1]
for config, P, E in (("aids-50+", 64, 256), ("n=60", 64, 128), ("n=70", 64, 64), ("n=80", 64, 64), ("n=100", 32, 64), ("n=120", 32, 32)):
    pairs = synthetic.make_pairs(config, P)
    edits = synthetic.make_edits(pairs, E)
    ps = PairSet.from_host(pairs, dev)
    ref = None
    for q in (1,):
        for k in ks:
            os.environ["GED_B200_WORK_QUEUE"] = str(q)
            os.environ["GED_B200_SMEM_WARPS"] = str(k)
            best = 1e9
            for rep in range(2):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                res = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2])
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            if ref is None:
                ref = res
            B = len(edits[0])
            same = bool(torch.equal(ref.ged, res.ged) and torch.equal(ref.rowmatch, res.rowmatch) and torch.equal(ref.iters, res.iters))
            print(f"{config} queue={q} smem_warps={k:2d}: B={B} {best:.1f} ms -> {B / best * 1e3:.0f} solves/s same={same}", flush=True)
