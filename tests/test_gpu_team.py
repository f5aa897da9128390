"""GPU: the CTA-per-candidate ("team") launches (ged_solve_desc.launch_hint = 2, csrc/ged_kernels.cu: ged_solve_team_kernel) --
all warps of a CTA share one candidate: LAP columns dealt out to the warps, rows of K.X dealt out to the warps.  They must
return the bits of the one-warp kernels and of the oracle: whole trajectories, every launch path (edits, scoring graph,
teacher forcing, Hungarian only, more candidates than one wave), every team width W = 2..7, and the helper-warp shape of pairs
of up to 32 nodes (one-warp LAP, K.X and update split over four warps)."""
import numpy as np
import pytest
import torch

from oracle import ged_oracle as go
from ppo_bihyb_b200 import solver, synthetic
from ppo_bihyb_b200.graphs import PairSet, label_node_cost

pytestmark = pytest.mark.gpu

TEAM, PACKED = 2, 0


def _same(a, b, trace=False):
    assert torch.equal(a.ged, b.ged) and torch.equal(a.ged_edited, b.ged_edited)
    assert torch.equal(a.rowmatch, b.rowmatch) and torch.equal(a.colmatch, b.colmatch)
    assert torch.equal(a.iters, b.iters) and torch.equal(a.status, b.status)
    if trace:
        for k in a.trace:
            x, y = a.trace[k], b.trace[k]
            assert torch.equal(x, y) or bool(((x == y) | (x.isnan() & y.isnan())).all()), k


@pytest.mark.parametrize("config,num", [("aids-20-30", 8), ("aids-30-50", 6), ("aids-50+", 5), ("n=100", 2)])
def test_team_trajectory_equals_one_warp_kernel_and_oracle(cuda_device, config, num):
    """Whole trajectories (Hungarian start, every iteration's cost matrix, projection, six scalars, iterate, iteration and
    LAP scan-step counts, result): team == one warp per candidate, and == the oracle on the first pairs."""
    pairs = synthetic.make_pairs(config, num, seed_offset=17)
    ps = PairSet.from_host(pairs, cuda_device)
    team = solver.solve_batch(ps, trace=True, dense_x=True, check=True, launch_hint=TEAM)
    one = solver.solve_batch(ps, trace=True, dense_x=True, check=True, launch_hint=PACKED)
    _same(team, one, trace=True)
    assert torch.equal(team.x, one.x)
    for b in range(min(2, num)):
        g1, g2 = pairs[b]
        o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), trace=True)
        n1, n2, T = g1.n, g2.n, o["iters"]
        R, C = n1 + 1, n2 + 1
        tr = team.trace
        assert int(team.iters[b]) == T and int(tr["lap_steps"][b]) == o["lap_steps"] and float(team.ged[b]) == o["ged"]
        sc = tr["scalars"][b, :T].cpu().numpy()
        for k, name in enumerate(("s_vv", "s_vb", "ub", "alpha", "beta", "t0")):
            assert np.array_equal(sc[:, k], o[name], equal_nan=True), name
        assert np.array_equal(tr["cost"][b, :T, :R * C].cpu().numpy().reshape(T, R, C), o["cost"])
        assert np.array_equal(tr["x_after"][b, :T, :R * C].cpu().numpy().reshape(T, R, C), o["x_after"])
        assert np.array_equal(tr["rowmatch"][b, :T, :n1].cpu().numpy(), o["b_rowmatch"])
        assert np.array_equal(team.rowmatch[b, :n1].cpu().numpy(), o["rowmatch"])


@pytest.mark.parametrize("n,iters", [(150, 4), (200, 3)])
def test_team_of_five_and_seven_warps(cuda_device, n, iters):
    pairs = synthetic.make_pairs(f"n={n}", 2, seed_offset=23)
    ps = PairSet.from_host(pairs, cuda_device)
    team = solver.solve_batch(ps, max_iter=iters, trace=True, check=True, launch_hint=TEAM)
    one = solver.solve_batch(ps, max_iter=iters, trace=True, check=True, launch_hint=PACKED)
    _same(team, one, trace=True)


def test_team_candidate_edits_size_classes_and_persistent_ctas(cuda_device):
    """A beam-step-like call (several size classes, edits scored on the base pair) and a call with more candidates than
    CTAs fit the device at once (persistent team CTAs pulling from the work queue)."""
    pairs = synthetic.make_pairs("aids-50+", 10, seed_offset=29)
    edits = synthetic.make_edits(pairs, 27, seed_offset=29)
    ps = PairSet.from_host(pairs, cuda_device)
    kw = dict(base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], check=True)
    team = solver.solve_batch(ps, launch_hint=TEAM, **kw)
    _same(team, solver.solve_batch(ps, launch_hint=PACKED, **kw))
    _same(team, solver.solve_batch(ps, **kw))                       # whatever the size-based default picks
    pairs = synthetic.make_pairs("aids-20-30", 10, seed_offset=30)
    edits = synthetic.make_edits(pairs, 27, seed_offset=30)
    ps = PairSet.from_host(pairs, cuda_device)
    kw = dict(base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], check=True)
    _same(solver.solve_batch(ps, launch_hint=TEAM, **kw), solver.solve_batch(ps, launch_hint=PACKED, **kw))
    pairs = synthetic.make_pairs("aids-30-50", 12, seed_offset=31)
    edits = synthetic.make_edits(pairs, 120, seed_offset=31)        # 1440 candidates
    ps = PairSet.from_host(pairs, cuda_device)
    kw = dict(base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], check=True)
    _same(solver.solve_batch(ps, launch_hint=TEAM, **kw), solver.solve_batch(ps, launch_hint=PACKED, **kw))


def test_team_scoring_graph_teacher_forcing_and_hungarian(cuda_device):
    pairs = synthetic.make_pairs("aids-30-50", 6, seed_offset=37)
    ps = PairSet.from_host(pairs, cuda_device)
    _same(solver.solve_batch(ps, solver="hungarian", check=True, launch_hint=TEAM),
          solver.solve_batch(ps, solver="hungarian", check=True, launch_hint=PACKED))
    # teacher forcing (x_init) from a mid-trajectory iterate of another run
    first = solver.solve_batch(ps, max_iter=5, trace=True, check=True, launch_hint=PACKED)
    x0 = first.trace["x_after"][:, 2].contiguous()
    a = solver.solve_batch(ps, max_iter=7, x_init=x0, trace=True, check=True, launch_hint=TEAM)
    b = solver.solve_batch(ps, max_iter=7, x_init=x0, trace=True, check=True, launch_hint=PACKED)
    _same(a, b, trace=True)
    # scoring on another graph 1 (GEDenv.step scores on the ORIGINAL pair after earlier edits, ged_env.py:122)
    edits = synthetic.make_edits(pairs, 5, seed_offset=37)
    score_adj, offs, o = [], [], 0
    for g1, _ in pairs:
        adj = g1.adj.copy()
        adj[0, 1] = adj[1, 0] = 1 - adj[0, 1]
        score_adj.append(adj.reshape(-1))
        offs.append(o)
        o += adj.size
    sa = torch.from_numpy(np.concatenate(score_adj)).to(cuda_device)
    so = torch.tensor(offs, dtype=torch.int64, device=cuda_device)
    kw = dict(base_idx=edits[0], edit_u=edits[1], edit_v=edits[2], score_adj1=sa, score_adj1_off=so, score_idx=edits[0], check=True)
    _same(solver.solve_batch(ps, launch_hint=TEAM, **kw), solver.solve_batch(ps, launch_hint=PACKED, **kw))


def test_team_status_bits(cuda_device):
    """A malformed graph (asymmetric adjacency) is reported per solve, as by the one-warp kernels."""
    from ppo_bihyb_b200.synthetic import HostGraph
    pairs = synthetic.make_pairs("aids-30-50", 3, seed_offset=41)
    bad = pairs[1][0].adj.copy()
    bad[0, 2], bad[2, 0] = 1, 0
    pairs[1] = (HostGraph(bad, pairs[1][0].labels), pairs[1][1])
    ps = PairSet.from_host(pairs, cuda_device)
    a = solver.solve_batch(ps, launch_hint=TEAM)
    b = solver.solve_batch(ps, launch_hint=PACKED)
    assert torch.equal(a.status, b.status) and int(a.status[1]) != 0 and int(a.status[0]) == 0
    assert torch.equal(a.ged[[0, 2]], b.ged[[0, 2]])


def test_team_explicit_node_costs_and_the_largest_pairs(cuda_device):
    """An explicit fp32 node-cost matrix (the route of the non-AIDS ``node_metric`` tables, ged_env.py:235-242) through the team
    launch, and pairs at the size limit (224 nodes), whose team CTA may not fit shared memory -- the call then takes the one-warp
    launch instead of failing."""
    rng = np.random.default_rng(7)
    pairs = synthetic.make_pairs("aids-30-50", 4, seed_offset=47)
    costs = []
    for g1, g2 in pairs:
        c = rng.integers(0, 4, (g1.n + 1, g2.n + 1)).astype(np.float32) / 2.0
        c[g1.n, g2.n] = 0.0
        costs.append(c)
    ps = PairSet.from_host(pairs, cuda_device, node_costs=costs)
    a = solver.solve_batch(ps, check=True, trace=True, max_iter=12, launch_hint=TEAM)
    b = solver.solve_batch(ps, check=True, trace=True, max_iter=12, launch_hint=PACKED)
    _same(a, b, trace=True)
    g1, g2 = pairs[0]
    o = go.solve(g1.adj, g2.adj, costs[0], max_iter=12)
    assert float(a.ged[0]) == o["ged"] and int(a.iters[0]) == o["iters"]
    big = synthetic.make_pairs("n=224", 1, seed_offset=49)
    psb = PairSet.from_host(big, cuda_device)
    _same(solver.solve_batch(psb, check=True, max_iter=2, launch_hint=TEAM), solver.solve_batch(psb, check=True, max_iter=2, launch_hint=PACKED))
