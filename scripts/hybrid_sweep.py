"""Development aid: solve throughput vs the number of shared-memory-backed warps per hybrid CTA (GED_B200_SMEM_WARPS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

dev = torch.device("cuda:0")
for config, P, E in (("n=70", 64, 64), ("n=80", 64, 64), ("n=60", 64, 64), ("n=100", 32, 64)):
    pairs = synthetic.make_pairs(config, P)
    edits = synthetic.make_edits(pairs, E)
    ps = PairSet.from_host(pairs, dev)
    ref = None
    for k in (-1, 0, 1, 2, 3, 4, 6):
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
        print(f"{config} smem_warps={k:2d}: B={B} {best:.1f} ms -> {B / best * 1e3:.0f} solves/s same={bool(torch.equal(ref.ged, res.ged))}", flush=True)
