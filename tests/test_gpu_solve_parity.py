"""GPU parity: the CUDA path (through the C-ABI) against the canonical-order CPU oracle, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import ged_oracle as go
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet, label_node_cost

pytestmark = pytest.mark.gpu


def _oracle_all(pairs, edits=None, **kw):
    out = []
    for b, (g1, g2) in enumerate(pairs if edits is None else [pairs[p] for p in edits[0]]):
        e = None if edits is None else (int(edits[1][b]), int(edits[2][b]))
        out.append(go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), edit=e, **kw))
    return out


@pytest.mark.parametrize("config,num", [("aids-20-30", 24), ("aids-30-50", 12), ("aids-50+", 6)])
def test_full_trajectory_bit_exact(cuda_device, config, num):
    pairs = synthetic.make_pairs(config, num, seed_offset=3)
    ps = PairSet.from_host(pairs, cuda_device)
    res = solver.solve_batch(ps, trace=True, dense_x=True, check=True)
    torch.cuda.synchronize()
    ora = _oracle_all(pairs, trace=True)
    for b, ((g1, g2), o) in enumerate(zip(pairs, ora)):
        n1, n2 = g1.n, g2.n
        R, C = n1 + 1, n2 + 1
        tr = res.trace
        assert np.array_equal(tr["init_cost"][b, :R * C].cpu().numpy().reshape(R, C), o["init_cost"]), f"init cost {b}"
        assert np.array_equal(tr["init_rowmatch"][b, :n1].cpu().numpy(), o["init_rowmatch"]), f"init assignment {b}"
        assert np.array_equal(tr["init_colmatch"][b, :n2].cpu().numpy(), o["init_colmatch"])
        T = o["iters"]
        assert int(res.iters[b]) == T, f"iteration count pair {b}"
        assert int(tr["lap_steps"][b]) == o["lap_steps"]
        sc = tr["scalars"][b, :T].cpu().numpy()
        for k, name in enumerate(("s_vv", "s_vb", "ub", "alpha", "beta", "t0")):
            assert np.array_equal(sc[:, k], o[name], equal_nan=True), f"{name} pair {b}"
        assert np.array_equal(tr["rowmatch"][b, :T, :n1].cpu().numpy(), o["b_rowmatch"])
        assert np.array_equal(tr["colmatch"][b, :T, :n2].cpu().numpy(), o["b_colmatch"])
        assert np.array_equal(tr["cost"][b, :T, :R * C].cpu().numpy().reshape(T, R, C), o["cost"]), f"cost pair {b}"
        assert np.array_equal(tr["x_after"][b, :T, :R * C].cpu().numpy().reshape(T, R, C), o["x_after"]), f"iterate pair {b}"
        assert np.array_equal(res.rowmatch[b, :n1].cpu().numpy(), o["rowmatch"])
        assert np.array_equal(res.colmatch[b, :n2].cpu().numpy(), o["colmatch"])
        assert float(res.ged[b]) == o["ged"]
        assert np.array_equal(res.x[b, :R * C].cpu().numpy().reshape(R, C), go.match_to_dense(o["rowmatch"], o["colmatch"]))


def test_candidate_edits_bit_exact(cuda_device):
    pairs = synthetic.make_pairs("aids-20-30", 6, seed_offset=5)
    edits = synthetic.make_edits(pairs, 8, seed_offset=1)
    ps = PairSet.from_host(pairs, cuda_device)
    res = solver.solve_batch(ps, base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], check=True)
    ora = _oracle_all(pairs, edits)
    ged = res.ged.cpu().numpy()
    ged_e = res.ged_edited.cpu().numpy()
    for b, o in enumerate(ora):
        g1, g2 = pairs[edits[0][b]]
        assert np.array_equal(res.rowmatch[b, :g1.n].cpu().numpy(), o["rowmatch"]), b
        assert ged[b] == o["ged"] and ged_e[b] == o["ged_edited"], b


def test_hungarian_solver_bit_exact(cuda_device):
    pairs = synthetic.make_pairs("aids-30-50", 10, seed_offset=9)
    ps = PairSet.from_host(pairs, cuda_device)
    res = solver.solve_batch(ps, solver="hungarian", check=True)
    for b, (g1, g2) in enumerate(pairs):
        o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), solver="hungarian")
        assert np.array_equal(res.rowmatch[b, :g1.n].cpu().numpy(), o["rowmatch"])
        assert np.array_equal(res.colmatch[b, :g2.n].cpu().numpy(), o["colmatch"])
        assert float(res.ged[b]) == o["ged"]
        assert int(res.iters[b]) == 0


def test_kv_matches_oracle_and_dense_k(cuda_device):
    rng = np.random.default_rng(11)
    pairs = synthetic.make_pairs("aids-20-30", 8, seed_offset=2)
    ps = PairSet.from_host(pairs, cuda_device)
    X = torch.zeros(len(pairs), ps.mat_stride)
    xs = []
    for b, (g1, g2) in enumerate(pairs):
        x = rng.random((g1.n + 1) * (g2.n + 1)).astype(np.float32)
        xs.append(x)
        X[b, :x.size] = torch.from_numpy(x)
    cost = solver.kv_batch(ps, X.to(cuda_device)).cpu().numpy()
    from ppo_bihyb_b200.graphs import dense_k_from_factors
    for b, (g1, g2) in enumerate(pairs):
        R, C = g1.n + 1, g2.n + 1
        Cn = label_node_cost(g1.labels, g2.labels)
        want = go.kv(g1.adj, g2.adj, Cn, xs[b].reshape(R, C))
        assert np.array_equal(cost[b, :R * C].reshape(R, C), want)
        K = dense_k_from_factors(g1.adj, g2.adj, Cn).double().numpy()
        dense = (K @ xs[b].astype(np.float64)).reshape(R, C)
        assert np.abs(dense - want).max() <= 1e-5 * np.abs(dense).max()     # fp32 tolerance of north_star: 1e-5 rel


def test_lsape_matches_scipy_order(cuda_device):
    """hungarian_lap stage: identical assignment on identical cost input, tie-heavy cases included."""
    rng = np.random.default_rng(5)
    mats, n1s, n2s = [], [], []
    for t in range(200):
        n1, n2 = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        kind = t % 4
        if kind == 0:
            c = rng.integers(0, 3, (n1 + 1, n2 + 1)).astype(np.float32)
        elif kind == 1:
            c = (rng.integers(0, 7, (n1 + 1, n2 + 1)) / 3.0).astype(np.float32)
        elif kind == 2:
            c = rng.random((n1 + 1, n2 + 1)).astype(np.float32)
        else:
            c = np.zeros((n1 + 1, n2 + 1), np.float32)
        mats.append(c)
        n1s.append(n1)
        n2s.append(n2)
    ms = max(m.size for m in mats)
    cost = torch.zeros(len(mats), ms)
    for b, m in enumerate(mats):
        cost[b, :m.size] = torch.from_numpy(m.reshape(-1))
    rm, cm, lb, st = solver.lsape_batch(cost.to(cuda_device), np.asarray(n1s), np.asarray(n2s), check=True)
    rm, cm, lb = rm.cpu().numpy(), cm.cpu().numpy(), lb.cpu().numpy()
    for b, m in enumerate(mats):
        r, c, want_lb, _ = go.lsape(m)
        assert np.array_equal(rm[b, :n1s[b]], r) and np.array_equal(cm[b, :n2s[b]], c), b
        assert lb[b] == np.float32(want_lb)
