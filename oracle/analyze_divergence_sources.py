"""TEST INFRASTRUCTURE ONLY -- which part of the arithmetic makes a free-running IPFP run leave the reference's?

    python oracle/analyze_divergence_sources.py [--pairs 200] [--config aids-20-30]       (build container: needs /root/reference)

The reference's ``ipfp_ged`` (``utils/ged_env.py:256-289``) is run unmodified; next to it, a restatement of the same loop is
run with ONE ingredient exchanged at a time:
  A  restatement with the reference's own operations (torch.mm for K v and for the scalars)   -> must equal the reference
  B  K v from the factorised canonical-order kernel arithmetic (oracle ``kv``), scalars as the reference computes them
  C  K v as the reference computes it, scalars in fp64 from the canonical dot products (what the kernels do)
  D  both exchanged (= the canonical order of oracle/ged_oracle.c, LAP through SciPy)
Every variant projects with the reference's own ``hungarian_lap`` (SciPy).  Printed: on how many pairs the final objective
equals the reference's.  Result (profiles/r02_parity_divergence.md): exchanging only the scalars (C) changes few pairs,
exchanging only the summation order of K v (B) changes as many as exchanging both -- the 1-ulp differences of ``cost`` at exact
ties decide, and no fp32 formulation of the scalars can win them back.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_variant(ref, k, n1, n2, kv_fn, fp64_scalars, max_iter=100):
    import torch
    v = ref.hungarian_ged(k, n1, n2)
    best, best_ub = v, float("inf")
    R, C = n1 + 1, n2 + 1
    for _ in range(max_iter):
        cost = kv_fn(v)
        b = ref.hungarian_lap(cost, n1, n2)[0]
        ub = ref.GEDenv.comp_ged(b, k)
        if ub < best_ub:
            best, best_ub = b, ub
        if fp64_scalars:
            c64, v64, b64 = cost.double(), v.double(), b.double()
            s_vv, s_vb = float((c64 * v64).sum()), float((c64 * b64).sum())
            alpha = s_vb - s_vv
            beta = (float(ub) - 2 * s_vb) + s_vv
            t0 = np.float32(-alpha / beta) if beta != 0 else np.float32(np.nan)
            jump = beta <= 0 or t0 >= 1
            gap = abs(s_vv - s_vb) / s_vv if s_vv != 0 else np.nan
            v_new = b if jump else v + torch.tensor(float(t0), dtype=torch.float32) * (b - v)
        else:
            alpha = torch.mm(torch.mm(v.view(1, -1), k), (b - v).view(-1, 1))
            beta = ref.GEDenv.comp_ged(b - v, k)
            t0 = -alpha / beta
            last = ref.GEDenv.comp_ged(v, k)
            gap = float(torch.abs(last - torch.mm(cost.view(1, -1), b.view(-1, 1))) / last)
            v_new = b if (beta <= 0 or t0 >= 1) else v + t0 * (b - v)
        v = v_new.reshape(R, C)
        if gap < 1e-3:
            break
    return float(ref.GEDenv.comp_ged(best, k))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200)
    ap.add_argument("--config", default="aids-20-30")
    args = ap.parse_args()
    import torch
    torch.set_num_threads(1)
    from oracle import ged_oracle as go, ref_shim
    from oracle.gen_golden import to_data
    from ppo_bihyb_b200 import synthetic
    from ppo_bihyb_b200.graphs import label_node_cost
    ref = ref_shim.load_reference()
    env = ref_shim.make_reference_env("ipfp", "AIDS-20-30", synthetic.NUM_LABELS)
    pairs = synthetic.make_pairs(args.config, args.pairs, seed_offset=31)
    same = dict(A=0, B=0, C=0, D=0)
    for g1, g2 in pairs:
        n1, n2 = g1.n, g2.n
        k = env.construct_k(to_data(g1), to_data(g2)).squeeze(0)
        want = float(ref.GEDenv.comp_ged(ref.ipfp_ged(k, n1, n2), k))
        Cn = label_node_cost(g1.labels, g2.labels)
        mm = lambda v: torch.mm(k, v.reshape(-1, 1)).reshape(n1 + 1, n2 + 1)
        canon = lambda v: torch.from_numpy(go.kv(g1.adj, g2.adj, Cn, v.numpy().reshape(n1 + 1, n2 + 1)))
        for name, kv_fn, f64 in (("A", mm, False), ("B", canon, False), ("C", mm, True), ("D", canon, True)):
            same[name] += int(run_variant(ref, k, n1, n2, kv_fn, f64) == want)
    print(f"{args.config}, {args.pairs} pairs: final objective identical to the unmodified reference on "
          f"A (restatement, reference ops) {same['A']}, B (canonical K v only) {same['B']}, "
          f"C (fp64 scalars only) {same['C']}, D (both) {same['D']}")


if __name__ == "__main__":
    main()
