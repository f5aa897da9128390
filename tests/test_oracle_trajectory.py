"""CPU: the oracle against the reference's OWN IPFP trajectory (fixtures of oracle/gen_golden_trajectory.py, made by running
the unmodified ``ipfp_ged``, reference ``utils/ged_env.py:256-289``).

Teacher forcing: every recorded iteration is replayed from the reference's iterate v_i, so the comparison does not inherit
the chaotic divergence of two free-running fp32 trajectories.  The aggregate test then bounds what that
divergence does to the objective, and where it starts.
"""
import numpy as np
import pytest

from oracle import ged_oracle as go
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.graphs import label_node_cost

from _golden import load, pairs_of

SHAPES = ["aids-20-30", "aids-30-50", "aids-50plus"]
REL = 1e-5            # north_star: intermediate iterates agree within 1e-5 relative error


def records_of(z):
    """Yield (pair index, iteration, v, cost, v_next, b_rowmatch, b_colmatch, ub, alpha, beta, t0)."""
    for r in range(len(z["pair"])):
        p = int(z["pair"][r])
        R, C = int(z["n1"][p]) + 1, int(z["n2"][p]) + 1
        sl = slice(z["mat_off"][r], z["mat_off"][r + 1])
        yield (p, int(z["iter"][r]), z["v"][sl].reshape(R, C), z["cost"][sl].reshape(R, C), z["v_next"][sl].reshape(R, C),
               z["b_rowmatch"][z["rm_off"][r]:z["rm_off"][r + 1]], z["b_colmatch"][z["cm_off"][r]:z["cm_off"][r + 1]],
               float(z["ub"][r]), float(z["alpha"][r]), float(z["beta"][r]), float(z["t0"][r]))


@pytest.mark.parametrize("name", SHAPES)
def test_teacher_forced_iterations_against_the_reference(name):
    """From the reference's own v_i: K v_i within 1e-5 relative of its ``torch.mm(k, v)`` (:269); the LSAPE of ITS cost matrix
    is its projection b_i, every tie included (:270, real fractional costs up to N = n1 + n2 ~ 190); ub exact; the
    line-search step (:275-282) reproduces its next iterate within 1e-5."""
    z = load("trajectory_" + name)
    pairs = pairs_of(z)
    worst_cost = worst_x = 0.0
    flipped = checked = 0
    for p, it, v, cost, v_next, brm, bcm, ub, alpha, beta, t0 in records_of(z):
        g1, g2 = pairs[p]
        Cn = label_node_cost(g1.labels, g2.labels)
        got = go.kv(g1.adj, g2.adj, Cn, v)
        err = float(np.abs(got - cost).max() / np.abs(cost).max())
        worst_cost = max(worst_cost, err)
        assert err <= REL, f"pair {p} iteration {it}: K v off by {err:.2e} relative"
        rm, cm, _, _ = go.lsape(cost)
        assert np.array_equal(rm, brm) and np.array_equal(cm, bcm), f"pair {p} iteration {it}: projection of the reference's cost differs"
        assert go.score(g1.adj, g2.adj, Cn, brm, bcm) == ub, f"pair {p} iteration {it}: ub"
        o = go.solve(g1.adj, g2.adj, Cn, max_iter=1, trace=True, x_init=v)
        checked += 1
        if not (np.array_equal(o["b_rowmatch"][0], brm) and np.array_equal(o["b_colmatch"][0], bcm)):
            # the oracle's own cost matrix (1e-7 away) tips a tie the other way: both projections must be optimal for the
            # reference's cost matrix to rounding
            b_or = go.match_to_dense(o["b_rowmatch"][0], o["b_colmatch"][0]).astype(np.float64)
            b_ref = go.match_to_dense(brm, bcm).astype(np.float64)
            c64 = cost.astype(np.float64)
            gap = ((c64 * b_or).sum() - (c64 * b_ref).sum()) / abs((c64 * b_ref).sum())
            assert abs(gap) <= 1e-6, f"pair {p} iteration {it}: projections differ by more than a rounding tie ({gap:.2e})"
            flipped += 1
            continue
        assert o["ub"][0] == ub
        dx = float(np.abs(o["x_after"][0] - v_next).max())           # entries live in [0, 1]: absolute = relative to the largest
        worst_x = max(worst_x, dx)
        assert dx <= REL, f"pair {p} iteration {it}: next iterate off by {dx:.2e} (t0 {o['t0'][0]} vs {t0})"
    print(f"{name}: {checked} teacher-forced iterations; max relative error of K v {worst_cost:.2e}, of the next iterate "
          f"{worst_x:.2e}; {flipped} iterations where the oracle's own cost matrix tips an exact tie the other way")
    assert flipped <= checked // 3          # exact ties recur along a trajectory (symmetric sub-structures): they are not rare events


@pytest.mark.parametrize("name,minimum", [("aids-20-30", 1000), ("aids-30-50", 200), ("aids-50plus", 64)])
def test_aggregate_objective_inside_the_reference_own_spread(name, minimum):
    """Free-running solves on >= 1000 / 200 / 64 pairs.  The reference is run at 1, 2, 4 and 8 MKL threads: the same code on
    the same inputs gives different objectives on a fraction of the pairs (summation order).  The canonical-order oracle
    is one more summation order; its paired mean difference to the 1-thread reference must be statistically
    indistinguishable from zero (|z| <= 3) -- as the reference's own thread variants are -- and every divergence must
    start at a tie decided by rounding while the iterates still agree within 1e-5."""
    z = load("aggregate_" + name)
    num, seed_offset = int(z["num"]), int(z["seed_offset"])
    assert num >= minimum
    pairs = synthetic.make_pairs(str(z["config"]), num, seed_offset)
    import hashlib
    h = hashlib.sha256()
    for g1, g2 in pairs:
        for g in (g1, g2):
            h.update(np.ascontiguousarray(g.adj).tobytes())
            h.update(np.ascontiguousarray(g.labels).tobytes())
    assert h.hexdigest() == str(z["checksum"]), "synthetic.make_pairs drifted: regenerate the aggregate fixtures"
    ged_or = np.asarray([go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels))["ged"] for g1, g2 in pairs], np.float32)
    assert np.array_equal(ged_or, z["ged_oracle"]), "the oracle's objectives changed since the fixture was generated"
    ref = z["ged_ref"]                      # [threads, pairs]
    threads = [int(t) for t in z["threads"]]
    base = ref[0].astype(np.float64)

    def paired(x):
        d = x.astype(np.float64) - base
        se = d.std(ddof=1) / np.sqrt(len(d)) if d.std() > 0 else 0.0
        return d.mean(), (d.mean() / se if se > 0 else 0.0), int((d == 0).sum())

    rows = [(f"reference, {t} threads", *paired(ref[k])) for k, t in enumerate(threads) if k > 0]
    m_or, z_or, same_or = paired(ged_or)
    for label, m, zz, same in rows + [("canonical-order oracle", m_or, z_or, same_or)]:
        print(f"{name}: {label:>24}: mean objective {base.mean() + m:9.4f} ({m:+.4f} vs 1 thread, z = {zz:+.2f}), identical on {same}/{num}")
    assert abs(z_or) <= 3.0, f"oracle's mean objective differs from the reference's beyond noise (z = {z_or:.2f})"
    assert abs(m_or) <= 0.01 * base.mean()
    # where the trajectories part
    kind, first_b = z["kind"], z["first_b"]
    diverged = kind > 0
    assert set(np.unique(kind)) <= {0, 1, 3}, "a divergence that did not start at a projection tie"
    gaps = np.maximum(np.abs(z["gap_ref"][kind == 1]), np.abs(z["gap_or"][kind == 1]))
    assert gaps.size == 0 or gaps.max() <= 1e-6, f"largest tie gap {gaps.max():.2e}"
    assert z["dv_before"].max() <= REL and z["dcost_before"].max() <= REL
    hist = np.bincount(first_b[kind == 1], minlength=1)
    print(f"{name}: trajectories identical to the end on {int((~diverged).sum())}/{num}; first differing projection at iteration "
          f"(median {int(np.median(first_b[kind == 1])) if gaps.size else -1}): " +
          ", ".join(f"{lo}-{hi}: {int(hist[lo:hi + 1].sum())}" for lo, hi in ((0, 0), (1, 1), (2, 4), (5, 9), (10, 24), (25, 49), (50, 99))) +
          f"; largest relative tie gap {gaps.max() if gaps.size else 0:.1e}; iterates agree within {z['dv_before'].max():.1e} until then")
