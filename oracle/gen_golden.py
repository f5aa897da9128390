"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference solver.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py

The reference (Thinklab-SJTU/PPO-BiHyb) ships no tests or known-answer vectors for this path,
so parity is pinned on outputs of the reference's own functions, imported through ``oracle/ref_shim.py`` and run
on CPU with ``torch.set_num_threads(1)`` (run-to-run deterministic at a fixed thread count).  Every fixture stores
its inputs, so the tests never need the reference tree (it does not exist on the GPU box).

Fixtures
  lsape.npz        hungarian_lap (ged_env.py:329-351 -> scipy linear_sum_assignment :407): cost -> pred_x, lb
  construct_k.npz  construct_k / node_metric (ged_env.py:160-243): tiny pairs -> dense K
  init.npz         heuristic_prediction_hun (ged_env.py:308-326): node_cost_mat by brute force, v0, lower bound
  iter0.npz        first IPFP iteration (ged_env.py:268-282) on the reference's dense K: cost, b, ub, alpha, beta, t0
  kv.npz           torch.mm(k, v) (ged_env.py:269) for random fractional v on the reference's dense K
  solve.npz        ipfp_ged / hungarian_ged + comp_ged end to end (ged_env.py:140-158): x, ged
  step.npz         GEDenv.step (ged_env.py:110-124): (pair, act) -> reward, new_solution, new edge_index
  other_metrics.npz  construct_k with the non-AIDS node_metric tables (ged_env.py:235-242) + hungarian / ipfp objectives
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from ppo_bihyb_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def to_data(g):
    return ref_shim.make_data(torch.from_numpy(g.features()), torch.from_numpy(g.edge_index()))


def x_to_match(x, n1, n2):
    """Dense (n1+1)x(n2+1) 0/1 -> (rowmatch [n1], colmatch [n2]) in the C-ABI's convention."""
    x = np.asarray(x)
    rowmatch = x[:n1].argmax(1).astype(np.int32)
    colmatch = np.where(x[n1, :n2] > 0, n1, x[:n1, :n2].argmax(0)).astype(np.int32)
    assert np.all(x[:n1].sum(1) == 1) and np.all(x[:, :n2].sum(0) == 1), "not an LSAPE solution"
    return rowmatch, colmatch


def pack_ragged(arrs, dtype):
    arrs = [np.asarray(a, dtype).reshape(-1) for a in arrs]
    off = np.zeros(len(arrs) + 1, np.int64)
    off[1:] = np.cumsum([a.size for a in arrs])
    return (np.concatenate(arrs) if arrs else np.zeros(0, dtype)), off


def pairs_blob(pairs, prefix=""):
    a1, o1 = pack_ragged([g1.adj for g1, _ in pairs], np.uint8)
    a2, o2 = pack_ragged([g2.adj for _, g2 in pairs], np.uint8)
    l1, lo1 = pack_ragged([g1.labels for g1, _ in pairs], np.int32)
    l2, lo2 = pack_ragged([g2.labels for _, g2 in pairs], np.int32)
    return {prefix + "n1": np.asarray([g1.n for g1, _ in pairs], np.int32),
            prefix + "n2": np.asarray([g2.n for _, g2 in pairs], np.int32),
            prefix + "adj1": a1, prefix + "adj1_off": o1, prefix + "adj2": a2, prefix + "adj2_off": o2,
            prefix + "lab1": l1, prefix + "lab1_off": lo1, prefix + "lab2": l2, prefix + "lab2_off": lo2}


def small_pairs(num, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return [(synthetic.molecule_like(int(rng.integers(lo, hi + 1)), rng),
             synthetic.molecule_like(int(rng.integers(lo, hi + 1)), rng)) for _ in range(num)]


def gen_lsape(ref):
    rng = np.random.default_rng(101)
    costs, xs, lbs = [], [], []
    shapes = []
    for t in range(96):
        n1, n2 = int(rng.integers(1, 26)), int(rng.integers(1, 26))
        kind = t % 6
        if kind == 0:
            c = rng.integers(0, 3, (n1 + 1, n2 + 1)).astype(np.float32)           # tie-heavy small integers
        elif kind == 1:
            c = (rng.integers(0, 7, (n1 + 1, n2 + 1)) / 3.0).astype(np.float32)   # k/3: inexact in binary
        elif kind == 2:
            c = rng.random((n1 + 1, n2 + 1)).astype(np.float32)
        elif kind == 3:
            c = np.zeros((n1 + 1, n2 + 1), np.float32)                            # everything ties
        elif kind == 4:
            c = rng.integers(0, 2, (n1 + 1, n2 + 1)).astype(np.float32)
            c[rng.random(c.shape) < 0.15] = np.inf                                # forbidden substitutions
            c[:, -1] = 1.0
            c[-1, :] = 1.0                                                        # deletions/insertions keep it feasible
        else:
            c = (rng.integers(-4, 5, (n1 + 1, n2 + 1)) * 0.25).astype(np.float32)  # negative entries
        x, lb = ref.hungarian_lap(torch.from_numpy(c), n1, n2)
        costs.append(c)
        xs.append(x.numpy())
        lbs.append(float(lb))
        shapes.append((n1, n2))
    cost, off = pack_ragged(costs, np.float32)
    x, _ = pack_ragged(xs, np.float32)
    np.savez_compressed(os.path.join(OUT, "lsape.npz"), n1=np.asarray([s[0] for s in shapes], np.int32),
                        n2=np.asarray([s[1] for s in shapes], np.int32), cost=cost, x=x, off=off,
                        lb=np.asarray(lbs, np.float64))


def gen_construct_k(env):
    pairs = small_pairs(6, 3, 7, 202)
    ks = []
    for g1, g2 in pairs:
        ks.append(env.construct_k(to_data(g1), to_data(g2)).squeeze(0).numpy())
    k, koff = pack_ragged(ks, np.float32)
    np.savez_compressed(os.path.join(OUT, "construct_k.npz"), k=k, k_off=koff, **pairs_blob(pairs))


def gen_init(ref, env):
    pairs = small_pairs(10, 4, 12, 303)
    ncs, xs, lbs = [], [], []
    for g1, g2 in pairs:
        n1, n2 = g1.n, g2.n
        k = env.construct_k(to_data(g1), to_data(g2)).squeeze(0)
        # node_cost_mat exactly as ged_env.py:310-315 builds it (the function returns only x and lb)
        k_prime = k.reshape(-1, n1 + 1, n2 + 1)
        nc = torch.empty(k_prime.shape[0])
        for i in range(k_prime.shape[0]):
            _, nc[i] = ref.hungarian_lap(k_prime[i], n1, n2)
        x, lb = ref.heuristic_prediction_hun(k, n1, n2)
        ncs.append(nc.reshape(n1 + 1, n2 + 1).numpy())
        xs.append(x.numpy())
        lbs.append(float(lb))
    nc, off = pack_ragged(ncs, np.float32)
    x, _ = pack_ragged(xs, np.float32)
    np.savez_compressed(os.path.join(OUT, "init.npz"), node_cost_mat=nc, x=x, off=off, lb=np.asarray(lbs),
                        **pairs_blob(pairs))


def gen_iter0_and_kv(ref, env):
    pairs = small_pairs(8, 8, 20, 404)
    rng = np.random.default_rng(405)
    rec = {k: [] for k in ("v0", "cost0", "b0", "ub0", "alpha0", "beta0", "t00", "v1", "vrand", "kv")}
    for g1, g2 in pairs:
        n1, n2 = g1.n, g2.n
        k = env.construct_k(to_data(g1), to_data(g2)).squeeze(0)
        v = ref.hungarian_ged(k, n1, n2)
        cost = torch.mm(k, v.view(-1, 1)).reshape(n1 + 1, n2 + 1)                      # ged_env.py:269
        b = ref.hungarian_lap(cost, n1, n2)[0]                                         # :270
        ub = ref.GEDenv.comp_ged(b, k)                                                 # :271
        alpha = torch.mm(torch.mm(v.view(1, -1), k), (b - v).view(-1, 1))              # :275
        beta = ref.GEDenv.comp_ged(b - v, k)                                           # :277
        t0 = -alpha / beta                                                             # :278
        v1 = b if (beta <= 0 or t0 >= 1) else v + t0 * (b - v)                         # :279-282
        vr = torch.from_numpy(rng.random((n1 + 1, n2 + 1)).astype(np.float32))
        kv = torch.mm(k, vr.view(-1, 1)).reshape(n1 + 1, n2 + 1)
        for name, val in (("v0", v), ("cost0", cost), ("b0", b), ("v1", v1), ("vrand", vr), ("kv", kv)):
            rec[name].append(val.numpy())
        rec["ub0"].append(float(ub))
        rec["alpha0"].append(float(alpha))
        rec["beta0"].append(float(beta))
        rec["t00"].append(float(t0))
    blob = {}
    for name in ("v0", "cost0", "b0", "v1", "vrand", "kv"):
        blob[name], blob["off"] = pack_ragged(rec[name], np.float32)
    for name in ("ub0", "alpha0", "beta0", "t00"):
        blob[name] = np.asarray(rec[name], np.float64)
    np.savez_compressed(os.path.join(OUT, "iter0.npz"), **blob, **pairs_blob(pairs))


def gen_solve(ref, env):
    """End-to-end reference solves.  These are the reference's OWN fp32/MKL trajectory: the canonical-order oracle
    matches them exactly only where the trajectory is order independent (parity contract); the tests
    assert exactness on hungarian and aggregate closeness + match rate on ipfp."""
    sets = [("aids-20-30", 40), ("aids-30-50", 16), ("aids-50+", 8)]
    for name, num in sets:
        pairs = synthetic.make_pairs(name, num, seed_offset=11)
        ged_ipfp, ged_hun, x_ipfp, x_hun, secs, ged_mt = [], [], [], [], [], []
        for g1, g2 in pairs:
            d1, d2 = to_data(g1), to_data(g2)
            t = time.time()
            env.solver_type = "ipfp"
            ged, k, _ = env.solve_feasible_ged(d1, d2, "ipfp")
            secs.append(time.time() - t)
            x = ref.ipfp_ged(k, g1.n, g2.n)
            assert float(ref.GEDenv.comp_ged(x, k)) == float(ged)
            xh = ref.hungarian_ged(k, g1.n, g2.n)
            # the same reference code with 4 MKL threads: shows the reference's own sensitivity to summation order
            torch.set_num_threads(4)
            ged_mt.append(float(ref.GEDenv.comp_ged(ref.ipfp_ged(k, g1.n, g2.n), k)))
            torch.set_num_threads(1)
            ged_ipfp.append(float(ged))
            ged_hun.append(float(ref.GEDenv.comp_ged(xh, k)))
            x_ipfp.append(x.numpy())
            x_hun.append(xh.numpy())
        xi, off = pack_ragged(x_ipfp, np.float32)
        xh, _ = pack_ragged(x_hun, np.float32)
        tag = name.replace("+", "plus")
        np.savez_compressed(os.path.join(OUT, f"solve_{tag}.npz"), ged_ipfp=np.asarray(ged_ipfp, np.float32),
                            ged_hungarian=np.asarray(ged_hun, np.float32),
                            ged_ipfp_4threads=np.asarray(ged_mt, np.float32), x_ipfp=xi, x_hungarian=xh, off=off,
                            ref_seconds=np.asarray(secs), **pairs_blob(pairs))
        print(f"  {name}: ipfp mean {np.mean(ged_ipfp):.3f} (4 threads: {np.mean(ged_mt):.3f}, identical on "
              f"{int(np.sum(np.asarray(ged_mt) == np.asarray(ged_ipfp)))}/{num})  hungarian mean {np.mean(ged_hun):.3f}  "
              f"{np.mean(secs):.3f} s/solve (1 thread)")


def gen_step(ref, env):
    pairs = synthetic.make_pairs("aids-20-30", 4, seed_offset=13)
    base, eu, ev = synthetic.make_edits(pairs, 6, seed_offset=3)
    env.solver_type = "ipfp"
    prev, rewards, sols, eis = [], [], [], []
    oris = {}
    for b in range(len(base)):
        g1, g2 = pairs[base[b]]
        d1, d2 = to_data(g1), to_data(g2)
        if base[b] not in oris:
            ged0, k0, _ = env.solve_feasible_ged(d1, d2, "hungarian")            # ori_greedy as generate_tuples builds it
            oris[base[b]] = (float(ged0), k0)
        ged0, k0 = oris[base[b]]
        act = torch.tensor([int(eu[b]), int(ev[b])])
        reward, new_g1, new_sol = env.step(d1, d2, k0, act, ged0)
        prev.append(ged0)
        rewards.append(float(reward))
        sols.append(float(new_sol))
        eis.append(new_g1.edge_index.numpy())
    ei, eoff = pack_ragged(eis, np.int64)
    np.savez_compressed(os.path.join(OUT, "step.npz"), base_idx=base, edit_u=eu, edit_v=ev,
                        prev_solution=np.asarray(prev, np.float32), reward=np.asarray(rewards, np.float32),
                        new_solution=np.asarray(sols, np.float32), new_edge_index=ei, new_edge_index_off=eoff,
                        **pairs_blob(pairs))


def gen_other_metrics(ref):
    """construct_k / node_metric for the NON-AIDS tables (ged_env.py:235-242: CMU / Willow [0, VERY_LARGE_INT, 0], any other
    dataset [0, 1, 0]; the whole feature row is compared, ori_feature_dim = 0) and hungarian_ged / ipfp_ged on that K."""
    rng = np.random.default_rng(606)
    out = {}
    for tag, dataset in (("willow", "Willow"), ("other", "SomeOtherDataset")):
        env = ref_shim.make_reference_env("ipfp", dataset, 0)
        pairs = small_pairs(4, 5, 11, 707 + len(tag))
        ks, xs, geds, gedh, feats1, feats2 = [], [], [], [], [], []
        for g1, g2 in pairs:
            f1 = np.eye(3, dtype=np.float32)[rng.integers(0, 3, g1.n)]
            f2 = np.eye(3, dtype=np.float32)[rng.integers(0, 3, g2.n)]
            d1 = ref_shim.make_data(torch.from_numpy(f1), torch.from_numpy(g1.edge_index()))
            d2 = ref_shim.make_data(torch.from_numpy(f2), torch.from_numpy(g2.edge_index()))
            k = env.construct_k(d1, d2).squeeze(0)
            xh = ref.hungarian_ged(k, g1.n, g2.n)
            ks.append(k.numpy())
            xs.append(xh.numpy())
            gedh.append(float(ref.GEDenv.comp_ged(xh, k)))
            geds.append(float(ref.GEDenv.comp_ged(ref.ipfp_ged(k, g1.n, g2.n), k)))
            feats1.append(f1)
            feats2.append(f2)
        out[tag + "_k"], out[tag + "_k_off"] = pack_ragged(ks, np.float32)
        out[tag + "_x_hungarian"], out[tag + "_off"] = pack_ragged(xs, np.float32)
        out[tag + "_feat1"], out[tag + "_f1_off"] = pack_ragged(feats1, np.float32)
        out[tag + "_feat2"], out[tag + "_f2_off"] = pack_ragged(feats2, np.float32)
        out[tag + "_ged_hungarian"] = np.asarray(gedh, np.float32)
        out[tag + "_ged_ipfp"] = np.asarray(geds, np.float32)
        out.update(pairs_blob(pairs, prefix=tag + "_"))
    np.savez_compressed(os.path.join(OUT, "other_metrics.npz"), **out)


def main():
    if "--other-metrics" in sys.argv:       # add this one fixture without regenerating the rest
        torch.set_num_threads(1)
        gen_other_metrics(ref_shim.load_reference())
        return
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not found; golden vectors can only be generated in the build container")
    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load_reference()
    env = ref_shim.make_reference_env("ipfp", "AIDS-20-30", synthetic.NUM_LABELS)
    import scipy
    t = time.time()
    for fn, args in ((gen_lsape, (ref,)), (gen_construct_k, (env,)), (gen_init, (ref, env)),
                     (gen_iter0_and_kv, (ref, env)), (gen_solve, (ref, env)), (gen_step, (ref, env)), (gen_other_metrics, (ref,))):
        print(fn.__name__, flush=True)
        fn(*args)
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write("Generated by oracle/gen_golden.py from the unmodified reference at /root/reference (utils/ged_env.py)\n"
                f"torch {torch.__version__}, scipy {scipy.__version__}, numpy {np.__version__}, torch threads 1, CPU fp32\n")
    print(f"done in {time.time() - t:.1f} s")


if __name__ == "__main__":
    main()
