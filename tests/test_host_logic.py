"""CPU: host-side logic of the drop-in layer (packing, edge toggles, factor recovery, sharding arithmetic)."""
import os
import numpy as np
import pytest
import torch

from ppo_bihyb_b200 import sharding, synthetic
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import (Data, PairSet, data_to_host, dense_k_from_factors, factors_from_dense_k,
                                   label_node_cost, node_cost_matrix)

from _golden import load, pairs_of


def test_synthetic_graphs_have_aids_shape():
    pairs = synthetic.make_pairs("aids-50+", 40)
    ns = np.asarray([g.n for p in pairs for g in p])
    assert ns.min() >= 50 and ns.max() <= 100 and 55 < ns.mean() < 65
    for g1, g2 in pairs[:8]:
        for g in (g1, g2):
            assert np.array_equal(g.adj, g.adj.T) and g.adj.diagonal().sum() == 0 and g.adj.max() == 1
            deg = g.adj.sum(1)
            assert deg.max() <= synthetic.MAX_DEGREE and 0.95 * g.n <= g.adj.sum() / 2 <= 1.2 * g.n
            assert g.features().shape == (g.n, 34)
    again = synthetic.make_pairs("aids-50+", 40)
    assert all(np.array_equal(a[0].adj, b[0].adj) and np.array_equal(a[1].labels, b[1].labels) for a, b in zip(pairs, again))


def test_pack_host_layout_is_the_c_struct_layout():
    pairs = synthetic.make_pairs("aids-20-30", 5)
    t, meta = PairSet.pack_host(pairs)
    assert meta["max_n1"] == max(g1.n for g1, _ in pairs) and meta["max_n2"] == max(g2.n for _, g2 in pairs)
    for p, (g1, g2) in enumerate(pairs):
        o = int(t["adj1_off"][p])
        assert np.array_equal(t["adj1"][o:o + g1.n ** 2].numpy().reshape(g1.n, g1.n), g1.adj)
        o = int(t["lab2_off"][p])
        assert np.array_equal(t["lab2"][o:o + g2.n].numpy(), g2.labels)
    assert t["adj1"].dtype == torch.uint8 and t["n1"].dtype == torch.int32 and t["adj1_off"].dtype == torch.int64


def test_data_round_trip_and_label_recovery():
    g = synthetic.make_pairs("aids-20-30", 1)[0][0]
    d = Data.from_host(g)
    adj, labels, feat = data_to_host(d, synthetic.NUM_LABELS)
    assert np.array_equal(adj, g.adj) and np.array_equal(labels, g.labels) and feat.shape[1] == 29
    # non one-hot features: labels cannot be recovered, the explicit node-cost matrix is used instead
    d.x = d.x.clone()
    d.x[0, :29] = 0
    _, labels, feat = data_to_host(d, synthetic.NUM_LABELS)
    assert labels is None
    cost = node_cost_matrix(feat, feat)
    assert cost.shape == (g.n + 1, g.n + 1) and cost[0, 0] == 0 and cost[-1, -1] == 0 and cost[0, -1] == 0


def test_toggle_edge_matches_reference_step_edge_lists():
    """GEDenv.step's edit (ged_env.py:111-121) against the reference's new edge_index, element for element."""
    z = load("step")
    pairs = pairs_of(z)
    for b in range(len(z["base_idx"])):
        g1 = pairs[z["base_idx"][b]][0]
        d = Data.from_host(g1)
        before = d.edge_index.clone()
        out = GEDenv.toggle_edge(d, torch.tensor([int(z["edit_u"][b]), int(z["edit_v"][b])]))
        want = z["new_edge_index"][z["new_edge_index_off"][b]:z["new_edge_index_off"][b + 1]].reshape(2, -1)
        assert np.array_equal(out.edge_index.numpy(), want)
        assert torch.equal(d.edge_index, before)                    # input untouched (reference clones, :111)
        data_to_host(d, synthetic.NUM_LABELS)                        # with a host view the edit is done on the host: same list
        out2 = GEDenv.toggle_edge(d, torch.tensor([int(z["edit_u"][b]), int(z["edit_v"][b])]))
        assert np.array_equal(out2.edge_index.numpy(), want) and torch.equal(d.edge_index, before)


def test_factor_recovery_from_dense_k():
    g1, g2 = synthetic.make_pairs("aids-20-30", 1)[0]
    Cn = label_node_cost(g1.labels, g2.labels)
    k = dense_k_from_factors(g1.adj, g2.adj, Cn)
    assert torch.equal(k, k.T)
    a1, a2, cost = factors_from_dense_k(k, g1.n, g2.n)
    assert np.array_equal(a1, g1.adj) and np.array_equal(a2, g2.adj) and np.array_equal(cost, Cn)


def test_unknown_solver_and_dataset_errors():
    with pytest.raises(AssertionError):
        GEDenv(solver_type="ga")                      # not in the reference's available_solvers either (ged_env.py:22-23)
    env = GEDenv("ipfp", "AIDS-20-30")
    g = Data.from_host(synthetic.make_pairs("aids-20-30", 1)[0][0])
    with pytest.raises(NotImplementedError):
        env.solve_feasible_ged(g, g, "nope")
    with pytest.raises(NotImplementedError):
        env.solve_feasible_ged(g, g, "beam")          # A*-beam: cache files only (dataset.generate_tuples)
    with pytest.raises(ValueError):
        GEDenv.comp_ged(torch.zeros(2), torch.zeros(2, 2))


def test_shard_indices_partition_and_balance():
    rng = np.random.default_rng(0)
    for B, world in ((0, 2), (5, 8), (4096, 8), (1000, 3)):
        n1, n2 = rng.integers(50, 101, B), rng.integers(50, 101, B)
        for cost in (None, sharding.work_estimate(n1, n2)):
            parts = [sharding.shard_indices(B, r, world, cost) for r in range(world)]
            allidx = np.concatenate(parts) if parts else np.zeros(0, np.int64)
            assert np.array_equal(np.sort(allidx), np.arange(B))
            assert max(len(p) for p in parts) <= -(-B // world) if B else True
        inter = [sharding.shard_indices(B, r, world, interleave=True) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate(inter)) if B else np.zeros(0), np.arange(B))
        if B >= 1000:
            w = sharding.work_estimate(n1, n2)
            loads = [w[sharding.shard_indices(B, r, world, w)].sum() for r in range(world)]
            assert max(loads) / (sum(loads) / world) < 1.02


def test_gather_results_single_process():
    local = torch.arange(6, dtype=torch.float32)
    idx = np.asarray([5, 0, 3, 1, 4, 2])
    out = sharding.gather_results(local, idx, 6)
    assert torch.equal(out[torch.as_tensor(idx)], local)


def test_host_view_cache_follows_edge_toggles():
    """GEDenv.toggle_edge carries the host view (adjacency counts) over to the edited graph; it must equal what a fresh
    conversion of the edited graph's tensors gives, through insertions, deletions and self-loop edits."""
    g = synthetic.make_pairs("aids-20-30", 1, seed_offset=9)[0][0]
    d = Data.from_host(g)
    data_to_host(d, synthetic.NUM_LABELS)                          # populates the cache
    rng = np.random.default_rng(1)
    for step in range(40):
        u, v = (int(t) for t in rng.integers(0, g.n, 2))
        d = GEDenv.toggle_edge(d, torch.tensor([u, v]))
        cached = data_to_host(d, synthetic.NUM_LABELS)
        fresh = Data(x=d.x, edge_index=d.edge_index)               # no cache on this one
        want = data_to_host(fresh, synthetic.NUM_LABELS)
        assert np.array_equal(cached[0], want[0]) and np.array_equal(cached[1], want[1]), (step, u, v)
    d.edge_index[0, 0] = (d.edge_index[0, 0] + 1) % g.n            # in-place modification invalidates the view
    assert np.array_equal(data_to_host(d, synthetic.NUM_LABELS)[0], data_to_host(Data(x=d.x, edge_index=d.edge_index), synthetic.NUM_LABELS)[0])


def test_cli_flags_and_yaml_overrides(tmp_path):
    """The drivers keep the reference's flag names / defaults (ged_ppo_bihyb_train.py:397-431) and its YAML semantics: every
    key overrides the flag of that name, an unknown key is an error (:440)."""
    import importlib
    cli = importlib.import_module("ppo_bihyb_b200.cli")
    a = cli.parse_arguments(["train"])
    assert (a.solver_type, a.dataset, a.gamma, a.k_epochs, a.update_timestep, a.max_timesteps) == ("hungarian", "AIDS700nef", 0.99, 4, 20, 300)
    assert (a.learning_rate, a.eps_clip, a.node_output_size, a.gnn_layers, a.search_size, a.batch_size) == (0.002, 0.2, 16, 10, 5, 1)
    assert a.betas == (0.9, 0.999) and a.lr_steps == [] and a.test_model_weight == "" and a.random_seed is None
    cfg = tmp_path / "c.yaml"
    cfg.write_text("solver_type: ipfp\ndataset: AIDS-30-50\nk_epochs: 10\nlr_steps: [100, 200]\nsearch_size: 3\n")
    b = cli.parse_arguments(["eval", "--config", str(cfg), "--k_epochs", "2"])
    assert (b.command, b.solver_type, b.dataset, b.k_epochs, b.lr_steps, b.search_size) == ("eval", "ipfp", "AIDS-30-50", 10, [100, 200], 3)
    cfg.write_text("no_such_key: 1\n")
    with pytest.raises(AssertionError, match="Unknown config key"):
        cli.parse_arguments(["train", "--config", str(cfg)])
    repo_cfg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "ged_b200.yaml")
    c = cli.parse_arguments(["train", "--config", repo_cfg])
    assert (c.solver_type, c.dataset, c.max_timesteps, c.batch_size) == ("ipfp", "AIDS-20-30", 10, 64)


def test_dense_k_factor_cache_does_not_alias_freed_tensors():
    """Regression: same-shape dense Ks built and freed in a loop get the same address and version from
    the allocator; the factor cache must never hand back an earlier K's factors."""
    from ppo_bihyb_b200.ged_env import _DenseKFactors
    rng = np.random.default_rng(5)
    n1, n2 = 6, 7
    _DenseKFactors._cache.clear()
    seen_ptrs = set()
    for rep in range(12):
        a1 = np.triu((rng.random((n1, n1)) < 0.4).astype(np.uint8), 1)
        a1 = a1 + a1.T
        a2 = np.triu((rng.random((n2, n2)) < 0.4).astype(np.uint8), 1)
        a2 = a2 + a2.T
        cost = label_node_cost(rng.integers(0, 3, n1).astype(np.int32), rng.integers(0, 3, n2).astype(np.int32))
        k = dense_k_from_factors(a1, a2, cost)
        seen_ptrs.add(k.data_ptr())
        fac = _DenseKFactors.get(k, n1, n2)
        assert fac is not None
        assert np.array_equal(fac[0].adj, a1) and np.array_equal(fac[1].adj, a2) and np.array_equal(fac[2], cost)
        again = _DenseKFactors.get(k, n1, n2)               # a hit on the live tensor
        assert again is fac
        del k
    assert len(_DenseKFactors._cache) <= _DenseKFactors._MAX
    # an in-place edit of a cached K is a different K (version counter) and must be re-derived
    k = dense_k_from_factors(a1, a2, cost)
    f0 = _DenseKFactors.get(k, n1, n2)
    k4 = k.view(n1 + 1, n2 + 1, n1 + 1, n2 + 1)
    was = float(k4[0, n2, 1, n2])
    k4[0, n2, 1, n2] = 1.0 - was            # breaks the symmetry of A1 -> not a construct_k matrix any more
    assert _DenseKFactors.get(k, n1, n2) is None and f0 is not None


def test_stored_instruction_counts_belong_to_the_built_library():
    """bench.py's roofline divides instruction counts stored in profiles/roofline_r02.json (ncu of the bench mix) by the step
    time it measures: the counts must have been taken on the library that is built from this tree."""
    import json
    from ppo_bihyb_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prof = json.load(open(os.path.join(root, "profiles", "roofline_r02.json")))
    version = _lib.load().ged_version().decode()
    for sig, w in prof["workloads"].items():
        assert w.get("library") == version, (sig, w.get("library"), version)
        assert w["warp_instructions_per_step"] == sum(c["warp_instructions"] for c in w["per_class"])
