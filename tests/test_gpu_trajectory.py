"""GPU: the CUDA path (through the C-ABI) teacher-forced along the REFERENCE's own IPFP trajectory, and the LAP kernel
against SciPy itself at the sizes the path runs at.

Fixtures: tests/golden/trajectory_*.npz -- every record is one iteration of the unmodified ``ipfp_ged`` (reference
``utils/ged_env.py:268-282``): its iterate v_i, its ``cost_i = torch.mm(k, v_i)``, its projection b_i, its scalars and its next
iterate (oracle/gen_golden_trajectory.py)."""
import numpy as np
import pytest
import torch

from oracle import ged_oracle as go
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet, label_node_cost

from _golden import load, pairs_of
from test_oracle_trajectory import REL, SHAPES, records_of

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hint", [0, 2], ids=["one-warp-per-candidate", "one-cta-per-candidate"])
@pytest.mark.parametrize("name", SHAPES)
def test_kernels_teacher_forced_along_the_reference_trajectory(cuda_device, name, hint):
    z = load("trajectory_" + name)
    pairs = pairs_of(z)
    recs = list(records_of(z))
    ps = PairSet.from_host(pairs, cuda_device)
    ms = ps.mat_stride
    B = len(recs)
    base = np.asarray([r[0] for r in recs], np.int32)
    n1 = np.asarray([pairs[p][0].n for p in base], np.int32)
    n2 = np.asarray([pairs[p][1].n for p in base], np.int32)
    X = torch.zeros(B, ms)
    COST = torch.zeros(B, ms)
    for b, (p, it, v, cost, v_next, *_rest) in enumerate(recs):
        X[b, :v.size] = torch.from_numpy(v.reshape(-1))
        COST[b, :cost.size] = torch.from_numpy(cost.reshape(-1))

    # (1) ged_lsape_batch on the reference's own cost matrices == the reference's projections, every tie included
    rm, cm, lb, st = solver.lsape_batch(COST.to(cuda_device), n1, n2, check=True)
    rm, cm = rm.cpu().numpy(), cm.cpu().numpy()
    for b, (p, it, v, cost, v_next, brm, bcm, *_s) in enumerate(recs):
        assert np.array_equal(rm[b, :n1[b]], brm) and np.array_equal(cm[b, :n2[b]], bcm), f"pair {p} iteration {it}: projection"

    # (2) one IPFP iteration from the reference's iterate (ged_solve_desc.x_init): K v within 1e-5 relative of the reference's
    #     mm, identical bits to the oracle's replay, ub exact, next iterate within 1e-5
    res = solver.solve_batch(ps, base_idx=base, x_init=X.to(cuda_device), max_iter=1, trace=True, check=True, launch_hint=hint)
    tr = {k: t.cpu().numpy() for k, t in res.trace.items()}
    worst_cost = worst_x = 0.0
    flipped = 0
    for b, (p, it, v, cost, v_next, brm, bcm, ub, alpha, beta, t0) in enumerate(recs):
        g1, g2 = pairs[p]
        R, C = g1.n + 1, g2.n + 1
        got = tr["cost"][b, 0, :R * C].reshape(R, C)
        err = float(np.abs(got - cost).max() / np.abs(cost).max())
        worst_cost = max(worst_cost, err)
        assert err <= REL, f"pair {p} iteration {it}: K v off by {err:.2e}"
        o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), max_iter=1, trace=True, x_init=v)
        assert np.array_equal(got, o["cost"][0])
        assert np.array_equal(tr["rowmatch"][b, 0, :g1.n], o["b_rowmatch"][0]) and np.array_equal(tr["colmatch"][b, 0, :g2.n], o["b_colmatch"][0])
        assert np.array_equal(tr["x_after"][b, 0, :R * C].reshape(R, C), o["x_after"][0])
        if not (np.array_equal(o["b_rowmatch"][0], brm) and np.array_equal(o["b_colmatch"][0], bcm)):
            flipped += 1                    # an exact tie tipped by the 1e-7 difference of the two cost matrices (CPU test bounds the gap)
            continue
        assert tr["scalars"][b, 0, 2] == ub
        dx = float(np.abs(tr["x_after"][b, 0, :R * C].reshape(R, C) - v_next).max())
        worst_x = max(worst_x, dx)
        assert dx <= REL, f"pair {p} iteration {it}: next iterate off by {dx:.2e}"
    print(f"{name}: {B} teacher-forced iterations on the GPU; K v within {worst_cost:.2e}, next iterate within {worst_x:.2e}, {flipped} tie flips")


def _lsape_cases(rng, count, lo, hi):
    mats = []
    for t in range(count):
        n1, n2 = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        kind = t % 3
        if kind == 0:
            c = rng.integers(0, 3, (n1 + 1, n2 + 1)).astype(np.float32)            # integer ties everywhere
        elif kind == 1:
            c = (rng.integers(0, 7, (n1 + 1, n2 + 1)) / 3.0).astype(np.float32)    # k/3: inexact in binary
        else:
            c = rng.random((n1 + 1, n2 + 1)).astype(np.float32)
        mats.append(c)
    return mats


def _expanded(c):
    """The (n1+n2)^2 matrix ``hungarian_lap`` hands to SciPy (reference ``utils/ged_env.py:333-341``)."""
    n1, n2 = c.shape[0] - 1, c.shape[1] - 1
    big = np.full((n1 + n2, n1 + n2), np.inf)
    big[:n1, :n2] = c[:n1, :n2]
    big[np.arange(n1), n2 + np.arange(n1)] = c[:n1, n2]
    big[n1 + np.arange(n2), np.arange(n2)] = c[n1, :n2]
    big[n1:, n2:] = 0.0
    return big


def _check_against_scipy(cuda_device, mats):
    from scipy.optimize import linear_sum_assignment
    n1s = np.asarray([m.shape[0] - 1 for m in mats], np.int32)
    n2s = np.asarray([m.shape[1] - 1 for m in mats], np.int32)
    ms = max(m.size for m in mats)
    cost = torch.zeros(len(mats), ms)
    for b, m in enumerate(mats):
        cost[b, :m.size] = torch.from_numpy(m.reshape(-1))
    rm, cm, _, _ = solver.lsape_batch(cost.to(cuda_device), n1s, n2s, check=True)
    rm, cm = rm.cpu().numpy(), cm.cpu().numpy()
    for b, m in enumerate(mats):
        n1, n2 = int(n1s[b]), int(n2s[b])
        row, col = linear_sum_assignment(_expanded(m))
        want_rm = np.where(col[:n1] < n2, col[:n1], n2)
        r4c = np.empty(n1 + n2, np.int64)
        r4c[col] = row
        want_cm = np.where(r4c[:n2] < n1, r4c[:n2], n1)
        assert np.array_equal(rm[b, :n1], want_rm) and np.array_equal(cm[b, :n2], want_cm), f"case {b} ({n1}x{n2}) differs from SciPy"
        orm, ocm, _, _ = go.lsape(m)
        assert np.array_equal(orm, want_rm) and np.array_equal(ocm, want_cm), f"case {b}: oracle differs from SciPy"


def test_lsape_kernel_equals_scipy_at_path_sizes(cuda_device):
    """SciPy itself (``scipy.optimize.linear_sum_assignment`` on the expanded matrix, reference ``utils/ged_env.py:333-343,407``)
    vs the kernel vs the oracle: 200 matrices with n1, n2 in [50, 100] (N = 100..200, slot counts S = 2..4) and 20 with
    n in [150, 200]."""
    rng = np.random.default_rng(2024)
    _check_against_scipy(cuda_device, _lsape_cases(rng, 200, 50, 100))
    _check_against_scipy(cuda_device, _lsape_cases(rng, 20, 150, 200))


def test_lsape_kernel_equals_scipy_on_real_ipfp_costs(cuda_device):
    """The same on the fractional cost matrices IPFP really produces (iterations 0..99 of AIDS-50+ solves, from the oracle)."""
    pairs = synthetic.make_pairs("aids-50+", 3, seed_offset=17)
    mats = []
    for g1, g2 in pairs:
        o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), trace=True)
        mats += [o["cost"][i] for i in range(0, o["iters"], 4)]
    _check_against_scipy(cuda_device, mats)


@pytest.mark.parametrize("n,num", [(150, 2), (200, 2)])
def test_full_solve_parity_at_large_sizes(cuda_device, n, num):
    """Full 100-iteration kernel-vs-oracle parity (objective, assignment, iteration and LAP scan-step counts) at the top of the
    BASELINE sweep (S = 5 and S = 7 slot classes, cost matrix in shared memory)."""
    pairs = synthetic.make_pairs(f"n={n}", num, seed_offset=23)
    ps = PairSet.from_host(pairs, cuda_device)
    res = solver.solve_batch(ps, check=True, lap_steps=True)
    steps = res.trace["lap_steps"].cpu().numpy()
    for b, (g1, g2) in enumerate(pairs):
        o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), trace=True)
        assert float(res.ged[b]) == o["ged"] and int(res.iters[b]) == o["iters"] and int(steps[b]) == o["lap_steps"]
        assert np.array_equal(res.rowmatch[b, :g1.n].cpu().numpy(), o["rowmatch"])
        assert np.array_equal(res.colmatch[b, :g2.n].cpu().numpy(), o["colmatch"])
