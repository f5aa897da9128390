"""TEST INFRASTRUCTURE ONLY -- whole-trajectory parity fixtures from the UNMODIFIED reference ``ipfp_ged``.

Run in the build container (needs /root/reference):   python oracle/gen_golden_trajectory.py [--quick]

What it records (reference ``utils/ged_env.py:256-289``; nothing in that file is edited -- ``hungarian_lap`` and
``GEDenv.comp_ged`` are wrapped for the duration of a call so that their arguments can be read):

  trajectory_<shape>.npz   TEACHER-FORCING fixtures, 8 pairs per shape (n~25 / n~40 / AIDS-50+): for a subsample of the
                           iterations i of the reference's own run -- the iterate v_i (:269 input), cost_i = mm(K, v_i)
                           (:269), b_i = hungarian_lap(cost_i) (:270), ub / alpha / beta / t0 (:271-278) and v_{i+1}
                           (:279-282).  The scalars are recomputed outside the loop with the reference's own expressions
                           at the same thread count; the recomputed v_{i+1} is asserted bit-equal to the recorded one.
  aggregate_<shape>.npz    the reference's objective on >= 1000 / 200 / 64 pairs at 1, 2, 4 and 8 MKL threads (the
                           reference's own sensitivity to summation order), and for the 1-thread run the iteration at which
                           the canonical-order oracle's projection first differs, with the size of the tie that decided it.

Pairs of the aggregate are regenerated from (config, count, seed_offset) by ``synthetic.make_pairs``; a checksum of the
packed adjacencies guards against generator drift.
"""
import argparse
import hashlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.environ.get("GED_GOLDEN_OUT", os.path.join(ROOT, "tests", "golden"))

SHAPES = [("aids-20-30", 1000, 21), ("aids-30-50", 200, 22), ("aids-50+", 64, 23)]       # (config, pairs, seed_offset)
THREADS = (1, 2, 4, 8)
TRAJ_PAIRS = 8
TRAJ_ITERS = (0, 1, 2, 3, 4, 6, 9, 14, 22, 35, 56, 80, 99)                                # subsample of the iterations kept

_ref = _env = _go = None


def _init_worker(threads):
    os.environ["MKL_DYNAMIC"] = "FALSE"
    os.environ["OMP_DYNAMIC"] = "FALSE"
    global _ref, _env, _go
    import torch
    torch.set_num_threads(threads)
    from oracle import ref_shim, ged_oracle
    from ppo_bihyb_b200 import synthetic
    _ref = ref_shim.load_reference()
    _env = ref_shim.make_reference_env("ipfp", "AIDS-20-30", synthetic.NUM_LABELS)
    _go = ged_oracle


def _to_data(g):
    import torch
    from oracle import ref_shim
    return ref_shim.make_data(torch.from_numpy(g.features()), torch.from_numpy(g.edge_index()))


def _match_of(x, n1, n2):
    x = np.asarray(x)
    rm = x[:n1].argmax(1).astype(np.int32)
    cm = np.where(x[n1, :n2] > 0, n1, x[:n1, :n2].argmax(0)).astype(np.int32)
    return rm, cm


def record_reference(g1, g2):
    """One unmodified ``ipfp_ged`` run with every iteration's (v_i, cost_i, b_i) read off the wrapped callees."""
    import torch
    ref, env = _ref, _env
    n1, n2 = g1.n, g2.n
    k = env.construct_k(_to_data(g1), _to_data(g2)).squeeze(0)
    lap_calls, ged_calls = [], []
    orig_lap, orig_comp = ref.hungarian_lap, ref.GEDenv.comp_ged

    def lap(cost, a, b):
        out = orig_lap(cost, a, b)
        lap_calls.append((cost, out[0]))
        return out

    def comp(_x, _k):
        ged_calls.append(_x)
        return orig_comp(_x, _k)

    ref.hungarian_lap, ref.GEDenv.comp_ged = lap, staticmethod(comp)
    try:
        x = ref.ipfp_ged(k, n1, n2)
    finally:
        ref.hungarian_lap, ref.GEDenv.comp_ged = orig_lap, staticmethod(orig_comp)
    init_calls = (n1 + 1) * (n2 + 1) + 1                     # heuristic_prediction_hun :312-313, :325
    iters = lap_calls[init_calls:]
    T = len(iters)
    assert len(ged_calls) == 3 * T, (len(ged_calls), T)
    v0 = lap_calls[init_calls - 1][1]
    vs = [ged_calls[3 * i + 2] for i in range(T)]            # last_v at :283 = v at the start of iteration i
    assert torch.equal(vs[0], v0)
    rec = dict(n1=n1, n2=n2, T=T, ged=float(orig_comp(x, k)), x=x.numpy(), v=[], cost=[], b=[], ub=[], alpha=[], beta=[], t0=[], v_next=[])
    for i in range(T):
        v, (cost, b) = vs[i], iters[i]
        ub = orig_comp(b, k)                                                           # :271
        alpha = torch.mm(torch.mm(v.view(1, -1), k), (b - v).view(-1, 1))              # :275
        beta = orig_comp(b - v, k)                                                     # :277
        t0 = -alpha / beta                                                             # :278
        v_next = b if (beta <= 0 or t0 >= 1) else v + t0 * (b - v)                     # :279-282
        if i + 1 < T:
            assert torch.equal(v_next.reshape(n1 + 1, n2 + 1), vs[i + 1].reshape(n1 + 1, n2 + 1)), f"iteration {i}: replayed update differs"
        rec["v"].append(v.numpy().reshape(n1 + 1, n2 + 1).copy())
        rec["cost"].append(cost.numpy().reshape(n1 + 1, n2 + 1).copy())
        rec["b"].append(b.numpy().reshape(n1 + 1, n2 + 1).copy())
        rec["v_next"].append(v_next.numpy().reshape(n1 + 1, n2 + 1).copy())
        for name, val in (("ub", ub), ("alpha", alpha), ("beta", beta), ("t0", t0)):
            rec[name].append(float(val))
    return rec


def first_divergence(g1, g2, rec):
    """Where the canonical-order oracle's run leaves the reference's (same pair), and how close the decision was."""
    from ppo_bihyb_b200.graphs import label_node_cost
    go = _go
    n1, n2 = g1.n, g2.n
    o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), trace=True)
    T = min(rec["T"], o["iters"])
    out = dict(ged_oracle=o["ged"], iters_oracle=o["iters"], first_b=-1, first_v=-1, gap_ref=np.nan, gap_or=np.nan,
               dv_before=0.0, dcost_before=0.0, kind=0)
    v_or = go.match_to_dense(o["init_rowmatch"], o["init_colmatch"])
    for i in range(T):
        dv = float(np.abs(v_or - rec["v"][i]).max())
        if dv > 1e-4 and out["first_v"] < 0:
            out["first_v"] = i
        b_or = go.match_to_dense(o["b_rowmatch"][i], o["b_colmatch"][i])
        scale = float(np.abs(rec["cost"][i]).max())
        if not np.array_equal(b_or, rec["b"][i]):
            out["first_b"] = i
            cr, co = rec["cost"][i].astype(np.float64), o["cost"][i].astype(np.float64)
            obj_ref = float((cr * rec["b"][i]).sum())
            # how much worse is the other side's projection on MY cost matrix, relative to the objective
            out["gap_ref"] = (float((cr * b_or).sum()) - obj_ref) / max(abs(obj_ref), 1e-30)
            obj_or = float((co * b_or).sum())
            out["gap_or"] = (float((co * rec["b"][i]).sum()) - obj_or) / max(abs(obj_or), 1e-30)
            out["dv_before"] = dv
            out["dcost_before"] = float(np.abs(o["cost"][i] - rec["cost"][i]).max()) / max(scale, 1e-30)
            # kind 1: the iterates still agreed (<= 1e-4) when the projections parted -> a tie decided by rounding
            # kind 2: the iterates had already parted (a line-search branch / t0 difference came first)
            out["kind"] = 1 if (out["first_v"] < 0 or out["first_v"] >= i) else 2
            break
        out["dv_before"] = max(out["dv_before"], dv)
        out["dcost_before"] = max(out["dcost_before"], float(np.abs(o["cost"][i] - rec["cost"][i]).max()) / max(scale, 1e-30))
        v_or = o["x_after"][i]
    if out["first_b"] < 0 and rec["T"] != o["iters"]:
        out["kind"] = 3                                         # same projections throughout, different stopping iteration
        out["first_b"] = T
    return out


def _job_traj(args):
    config, seed_offset, index = args
    from ppo_bihyb_b200 import synthetic
    g1, g2 = synthetic.make_pairs(config, index + 1, seed_offset)[index]
    rec = record_reference(g1, g2)
    div = first_divergence(g1, g2, rec)
    return index, rec, div


def _job_ged(args):
    config, seed_offset, lo, hi = args
    from ppo_bihyb_b200 import synthetic
    pairs = synthetic.make_pairs(config, hi, seed_offset)[lo:hi]
    out = []
    for g1, g2 in pairs:
        ged, _, _ = _env.solve_feasible_ged(_to_data(g1), _to_data(g2), "ipfp")
        out.append(float(ged))
    return lo, out


def _job_div(args):
    """1-thread reference run with first-divergence analysis, for a block of pairs (trajectories are not kept)."""
    config, seed_offset, lo, hi = args
    from ppo_bihyb_b200 import synthetic
    pairs = synthetic.make_pairs(config, hi, seed_offset)[lo:hi]
    out = []
    for g1, g2 in pairs:
        rec = record_reference(g1, g2)
        div = first_divergence(g1, g2, rec)
        out.append((rec["ged"], rec["T"], div))
    return lo, out


def pairs_checksum(pairs):
    h = hashlib.sha256()
    for g1, g2 in pairs:
        for g in (g1, g2):
            h.update(np.ascontiguousarray(g.adj).tobytes())
            h.update(np.ascontiguousarray(g.labels).tobytes())
    return h.hexdigest()


def pack(arrs, dtype):
    arrs = [np.asarray(a, dtype).reshape(-1) for a in arrs]
    off = np.zeros(len(arrs) + 1, np.int64)
    off[1:] = np.cumsum([a.size for a in arrs])
    return (np.concatenate(arrs) if arrs else np.zeros(0, dtype)), off


def gen_trajectories(config, seed_offset, cores):
    from oracle.gen_golden import pairs_blob
    from ppo_bihyb_b200 import synthetic
    pairs = synthetic.make_pairs(config, TRAJ_PAIRS, seed_offset)
    with mp.get_context("spawn").Pool(min(cores, TRAJ_PAIRS), initializer=_init_worker, initargs=(1,)) as pool:
        res = sorted(pool.map(_job_traj, [(config, seed_offset, i) for i in range(TRAJ_PAIRS)]), key=lambda r: r[0])
    blob = {k: [] for k in ("v", "cost", "v_next")}
    rows = dict(pair=[], iter=[], ub=[], alpha=[], beta=[], t0=[])
    brm, bcm = [], []
    for p, rec, div in res:
        keep = sorted({i for i in TRAJ_ITERS if i < rec["T"]} | {rec["T"] - 1})
        for i in keep:
            rows["pair"].append(p)
            rows["iter"].append(i)
            for name in ("ub", "alpha", "beta", "t0"):
                rows[name].append(rec[name][i])
            for name in ("v", "cost", "v_next"):
                blob[name].append(rec[name][i])
            rm, cm = _match_of(rec["b"][i], rec["n1"], rec["n2"])
            brm.append(rm)
            bcm.append(cm)
    out = {}
    for name in ("v", "cost", "v_next"):
        out[name], out["mat_off"] = pack(blob[name], np.float32)
    out["b_rowmatch"], out["rm_off"] = pack(brm, np.int32)
    out["b_colmatch"], out["cm_off"] = pack(bcm, np.int32)
    out["pair"] = np.asarray(rows["pair"], np.int32)
    out["iter"] = np.asarray(rows["iter"], np.int32)
    for name in ("ub", "alpha", "beta", "t0"):
        out[name] = np.asarray(rows[name], np.float64)            # fp32 values of the reference, widened
    out["ref_iters"] = np.asarray([r[1]["T"] for r in res], np.int32)
    out["ref_ged"] = np.asarray([r[1]["ged"] for r in res], np.float32)
    x, xoff = pack([r[1]["x"] for r in res], np.float32)
    out["ref_x"], out["x_off"] = x, xoff
    tag = config.replace("+", "plus")
    np.savez_compressed(os.path.join(OUT, f"trajectory_{tag}.npz"), **out, **pairs_blob(pairs))
    print(f"  trajectory_{tag}: {len(out['pair'])} teacher-forcing records from {TRAJ_PAIRS} pairs, "
          f"iterations run {out['ref_iters'].tolist()}", flush=True)


def gen_aggregate(config, num, seed_offset, cores):
    from ppo_bihyb_b200 import synthetic
    pairs = synthetic.make_pairs(config, num, seed_offset)
    block = max(1, num // (4 * cores))
    blocks = [(config, seed_offset, lo, min(num, lo + block)) for lo in range(0, num, block)]
    ged = {}
    t_start = time.time()
    with mp.get_context("spawn").Pool(cores, initializer=_init_worker, initargs=(1,)) as pool:
        res = sorted(pool.map(_job_div, blocks), key=lambda r: r[0])
    flat = [r for _, part in res for r in part]
    ged[1] = np.asarray([r[0] for r in flat], np.float32)
    secs1 = time.time() - t_start
    ref_iters = np.asarray([r[1] for r in flat], np.int32)
    div = {k: np.asarray([r[2][k] for r in flat]) for k in flat[0][2]}
    for t in THREADS[1:]:
        procs = max(1, cores // t)
        with mp.get_context("spawn").Pool(procs, initializer=_init_worker, initargs=(t,)) as pool:
            res = sorted(pool.map(_job_ged, blocks), key=lambda r: r[0])
        ged[t] = np.asarray([g for _, part in res for g in part], np.float32)
    tag = config.replace("+", "plus")
    np.savez_compressed(os.path.join(OUT, f"aggregate_{tag}.npz"), config=config, num=num, seed_offset=seed_offset,
                        checksum=pairs_checksum(pairs), threads=np.asarray(THREADS, np.int32),
                        ged_ref=np.stack([ged[t] for t in THREADS]), ref_iters=ref_iters,
                        ged_oracle=div["ged_oracle"].astype(np.float32), iters_oracle=div["iters_oracle"].astype(np.int32),
                        first_b=div["first_b"].astype(np.int32), first_v=div["first_v"].astype(np.int32),
                        gap_ref=div["gap_ref"].astype(np.float64), gap_or=div["gap_or"].astype(np.float64),
                        dv_before=div["dv_before"].astype(np.float64), dcost_before=div["dcost_before"].astype(np.float64),
                        kind=div["kind"].astype(np.int32))
    same = {t: int((ged[t] == ged[1]).sum()) for t in THREADS}
    print(f"  aggregate_{tag}: {num} pairs, reference means " + ", ".join(f"{t}t {ged[t].mean():.4f}" for t in THREADS) +
          f"; identical to 1 thread: {same}; oracle mean {div['ged_oracle'].mean():.4f}, identical on "
          f"{int((div['ged_oracle'].astype(np.float32) == ged[1]).sum())}; 1-thread pass {secs1:.0f} s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="tiny counts (smoke run of the generator itself)")
    ap.add_argument("--only", choices=["trajectory", "aggregate"], default=None)
    ap.add_argument("--cores", type=int, default=os.cpu_count())
    args = ap.parse_args()
    from oracle import ref_shim
    if not ref_shim.reference_available():
        raise SystemExit("reference tree not found; golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    t = time.time()
    for config, num, seed_offset in SHAPES:
        if args.quick:
            num = 8
        if args.only in (None, "trajectory"):
            gen_trajectories(config, seed_offset + 100, args.cores)
        if args.only in (None, "aggregate"):
            gen_aggregate(config, num, seed_offset, args.cores)
    print(f"done in {time.time() - t:.0f} s")


if __name__ == "__main__":
    main()
