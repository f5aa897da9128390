"""Row f-4, GPU part: RRWM / GA on the real kernels (``ged_kv_batch``, ``ged_lsape_batch``, the Hungarian
start) against goldens of the UNMODIFIED reference ``rrwm_ged`` / ``ga_ged`` and through the drop-in ``ged_env`` entry points."""
import numpy as np
import pytest
import torch

from ppo_bihyb_b200 import baselines, ged_env, solver
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data, PairSet, dense_k_from_factors, label_node_cost

from _golden import load, pairs_of
from test_baselines import RTOL_EARLY, RTOL_LATE, _close, _mats

pytestmark = pytest.mark.gpu


def test_rrwm_iterates_and_projection_match_reference(cuda_device):
    z = load("baselines")
    pairs = pairs_of(z)
    ps = PairSet.from_host(pairs, cuda_device)
    for key, it, rtol in (("rrwm_v1", 1, RTOL_EARLY), ("rrwm_v5", 5, 10 * RTOL_EARLY), ("rrwm_v", 100, RTOL_LATE)):
        tr = {}
        rm, cm, iters = baselines.rrwm_batch(ps, max_iter=it, trace=tr)
        v = tr["v"].cpu().numpy()
        for b, want in enumerate(_mats(z, key)):
            got = v[b, :want.shape[0], :want.shape[1]]
            assert _close(got, want, rtol), (key, b, np.abs(got - want).max() / np.abs(want).max())
            assert abs(float(got.sum()) - 1.0) < 1e-4                                  # L1-normalised (:470-471)
        assert float(np.abs(v[~baselines._Ragged(ps).valid.cpu().numpy()]).sum()) == 0.0    # padding stays zero
    ged = solver.score_batch(ps, rm, cm).cpu().numpy()
    same = 0
    for b, want in enumerate(_mats(z, "rrwm_x")):
        x = solver.match_to_dense(rm[b], cm[b], want.shape[0] - 1, want.shape[1] - 1).cpu().numpy()
        assert np.all(x[:-1].sum(1) == 1) and np.all(x[:, :-1].sum(0) == 1)             # a feasible edit path
        if np.array_equal(x, want):
            same += 1
            assert ged[b] == z["rrwm_ged"][b]
    assert same >= len(pairs) - 1


def test_ga_matches_reference(cuda_device):
    z = load("baselines")
    pairs = pairs_of(z)
    ps = PairSet.from_host(pairs, cuda_device)
    tr = {}
    rm, cm = baselines.ga_batch(ps, trace=tr)
    ged = solver.score_batch(ps, rm, cm).cpu().numpy()
    soft = tr["kv_soft"].cpu().numpy()
    same = close = 0
    for b, (want_soft, want_x) in enumerate(zip(_mats(z, "ga_soft"), _mats(z, "ga_x"))):
        close += bool(_close(soft[b, :want_x.shape[0], :want_x.shape[1]], want_soft, RTOL_EARLY))
        x = solver.match_to_dense(rm[b], cm[b], want_x.shape[0] - 1, want_x.shape[1] - 1).cpu().numpy()
        if np.array_equal(x, want_x):
            same += 1
            assert ged[b] == z["ga_ged"][b]
    assert close >= len(pairs) - 2 and same >= len(pairs) - 1       # threshold-sitting pairs: see tests/test_baselines.py


def test_drop_in_entry_points(cuda_device):
    """``rrwm_ged(k, n1, n2)`` on a FactoredK and on the dense K, ``solve_feasible_ged(.., 'rrwm')`` and an rrwm ``step``."""
    z = load("baselines")
    pairs = pairs_of(z)
    env = GEDenv("rrwm", "AIDS-20-30", device=cuda_device)
    wants = _mats(z, "rrwm_x")
    hits = 0
    for b in (0, 2, 5):
        g1, g2 = pairs[b]
        d1, d2 = Data.from_host(g1, cuda_device), Data.from_host(g2, cuda_device)
        fk = env.construct_k(d1, d2)
        x_f = ged_env.rrwm_ged(fk, g1.n, g2.n)
        dense = dense_k_from_factors(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels)).to(cuda_device)
        x_d = ged_env.rrwm_ged(dense, torch.tensor([g1.n]), torch.tensor([g2.n]))
        assert x_f.shape == (g1.n + 1, g2.n + 1) and torch.equal(x_f, x_d)
        sol, k, sec = env.solve_feasible_ged(d1, d2, "rrwm")
        assert sol.shape == (1,) and sec > 0 and float(sol) == float(env.comp_ged(x_f, fk))
        if np.array_equal(x_f.cpu().numpy(), wants[b]):
            hits += 1
            assert float(sol) == float(z["rrwm_ged"][b])
        x_g = ged_env.ga_ged(fk, g1.n, g2.n)
        assert np.array_equal(x_g.cpu().numpy(), _mats(z, "ga_x")[b])
    assert hits >= 2
    # step with the rrwm solver: reward = previous - new solution, new solution scored on the ORIGINAL K (ged_env.py:122,158)
    g1, g2 = pairs[4]
    d1, d2 = Data.from_host(g1, cuda_device), Data.from_host(g2, cuda_device)
    prev, ori_k, _ = env.solve_feasible_ged(d1, d2, "rrwm")
    act = torch.tensor([0, 3])
    reward, new_g, new_sol = env.step(d1, d2, ori_k, act, prev)
    assert new_g.edge_index.shape[1] in (d1.edge_index.shape[1] + 2, d1.edge_index.shape[1] - 2)
    x_new = ged_env.rrwm_ged(env.construct_k(new_g, d2), g1.n, g2.n)
    assert float(new_sol) == float(env.comp_ged(x_new, ori_k)) and float(reward) == float(prev) - float(new_sol)
    with pytest.raises(NotImplementedError):
        env.solve_feasible_ged(d1, d2, "beam")


def test_sinkhorn_kernel_matches_the_torch_statement(cuda_device):
    """``ged_sinkhorn_batch`` (all passes in shared memory) against the batched torch statement of utils/sinkhorn.py:143-239
    that the CPU tests pin on the reference's goldens: ragged sizes, per-matrix temperatures, 1 / 10 / 100 passes."""
    rng = np.random.default_rng(17)
    B, Rm, Cm = 9, 37, 41
    rows = torch.tensor([37, 1, 5, 20, 36, 37, 2, 30, 11], dtype=torch.int32)
    cols = torch.tensor([41, 1, 41, 3, 40, 7, 2, 30, 29], dtype=torch.int32)
    s = torch.from_numpy(rng.normal(size=(B, Rm, Cm)).astype(np.float32)).to(cuda_device)
    tau = torch.from_numpy(rng.uniform(0.05, 2.0, size=B).astype(np.float32)).to(cuda_device)
    valid = (torch.arange(Rm)[None, :, None] < rows[:, None, None]) & (torch.arange(Cm)[None, None, :] < cols[:, None, None])
    valid = valid.to(cuda_device)
    for iters, tol in ((1, 2e-6), (10, 1e-5), (100, 1e-4)):
        got = solver.sinkhorn_batch(s, rows, cols, tau, iters)
        want = baselines._sinkhorn_log(s, valid, tau[:, None, None], iters)
        assert got.shape == want.shape and float(got[~valid].abs().sum()) == 0.0
        err = (got - want).abs().amax((1, 2)) / want.abs().amax((1, 2))
        assert float(err.max()) <= tol, (iters, err)
        if iters % 2 == 0:           # the last pass normalised the columns
            sums = got.sum(1)
            ok = torch.arange(Cm, device=cuda_device)[None, :] < cols.to(cuda_device)[:, None]
            assert float((sums[ok] - 1).abs().max()) < 1e-4
    # zero passes: exp(s / tau)
    assert torch.allclose(solver.sinkhorn_batch(s, rows, cols, tau, 0)[valid], torch.exp(s / tau[:, None, None])[valid], rtol=1e-6)
