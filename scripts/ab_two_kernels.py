"""Experiment (development aid): does a 5th, shared-memory-backed CTA per SM next to the four tensor-memory CTAs pay for the
S = 2 class?  The candidate list is split statically between a tensor-memory launch and a shared-memory launch (4-warp CTAs) on
two streams; needs a build with -DGED_TMEM_CTAS_S2=5 (96 registers) so that five CTAs fit the register file.
    GED_B200_LIBRARY=.../libged_b200.lb5.so python scripts/ab_two_kernels.py [n=60] [P] [E]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet

config = sys.argv[1] if len(sys.argv) > 1 else "n=60"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
E = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda:0")
pairs = synthetic.make_pairs(config, P, seed_offset=3)
base, eu, ev = synthetic.make_edits(pairs, E, seed_offset=3)
ps = PairSet.from_host(pairs, dev)
B = len(base)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def run(frac):
    idx = np.arange(B)
    to_smem = (idx % 100) < int(round(frac * 100))
    parts = [(idx[~to_smem], {}), (idx[to_smem], {"GED_B200_COST_STORE": "0", "GED_B200_SMEM_CTA_WARPS": "4"})]
    best, chk = None, None
    for rep in range(3):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = []
        for (sel, env), st in zip(parts, (s1, s2)):
            if len(sel) == 0:
                continue
            for k, v in env.items():
                os.environ[k] = v
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                outs.append((sel, solver.solve_batch(ps, base_idx=base[sel], edit_u=eu[sel], edit_v=ev[sel])))
            for k in env:
                os.environ.pop(k)
        for st in (s1, s2):
            torch.cuda.current_stream(dev).wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if rep > 0:
            best = ms if best is None else min(best, ms)
        ged = np.zeros(B, np.float64)
        for sel, res in outs:
            assert int((res.status != 0).sum()) == 0
            ged[sel] = res.ged.double().cpu().numpy()
        chk = ged.sum()
    return best, chk


for frac in (0.0, 0.1, 0.15, 0.2, 0.25, 0.3):
    ms, chk = run(frac)
    print(f"{config} B={B}: {int(frac * 100):2d} % of the candidates to the shared-memory launch: {ms:.1f} ms -> {B / ms * 1e3:,.0f} solves/s (checksum {chk:.0f})", flush=True)
