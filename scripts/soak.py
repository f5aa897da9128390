"""Soak test (development aid; the pytest suite holds the bounded version): many random candidates per size class, CUDA path
against oracle/ged_oracle.c, bit for bit -- objective, assignment, iteration count.  Usage: soak.py [candidates per config]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import torch
from oracle import ged_oracle as go
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet, label_node_cost

per = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
go.lib()
threads = os.cpu_count() or 1
bad_total = 0
CONFIGS = (("aids-20-30", 32), ("aids-30-50", 32), ("aids-50+", 48), ("n=70", 8), ("n=80", 8), ("n=85", 8), ("n=100", 6), ("n=130", 3))
only = os.environ.get("SOAK_CONFIGS")        # e.g. SOAK_CONFIGS=aids-50+,n=80,n=85
for config, P in (c for c in CONFIGS if only is None or c[0] in only.split(",")):
    pairs = synthetic.make_pairs(config, P, seed_offset=4242)
    E = max(1, per // P)
    edits = synthetic.make_edits(pairs, E, seed_offset=4242)
    ps = PairSet.from_host(pairs, dev)
    t0 = time.time()
    res = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], check=True)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    ged, it = res.ged.cpu().numpy(), res.iters.cpu().numpy()
    rm, cm = res.rowmatch.cpu().numpy(), res.colmatch.cpu().numpy()

    def one(b):
        g1, g2 = pairs[edits[0][b]]
        o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), edit=(int(edits[1][b]), int(edits[2][b])))
        return (o["ged"] == ged[b] and o["iters"] == it[b] and np.array_equal(o["rowmatch"], rm[b, :g1.n])
                and np.array_equal(o["colmatch"], cm[b, :g2.n]))
    t0 = time.time()
    with ThreadPoolExecutor(threads) as ex:
        ok = list(ex.map(one, range(len(edits[0]))))
    bad = len(ok) - sum(ok)
    bad_total += bad
    print(f"{config}: {len(ok)} candidates, {bad} mismatches; GPU {t_gpu:.2f} s, oracle {time.time() - t0:.1f} s on {threads} threads", flush=True)
print("SOAK", "FAILED" if bad_total else "OK")
sys.exit(1 if bad_total else 0)
