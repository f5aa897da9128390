"""GPU: the device-resident environment state (ppo_bihyb_b200.device_env) against the list-of-Data path it shortcuts
(GEDenv.step_batch / rollout_episode), which the other GPU tests pin on the reference goldens and the oracle."""
import copy

import numpy as np
import pytest
import torch

from ppo_bihyb_b200 import ppo as ppo_mod
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.device_env import DeviceGEDState
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data

pytestmark = pytest.mark.gpu


def _pairs(cuda_device, config="aids-20-30", num=12, seed=901):
    pairs = synthetic.make_pairs(config, num, seed_offset=seed)
    return ([Data.from_host(a, cuda_device) for a, _ in pairs], [Data.from_host(b, cuda_device) for _, b in pairs])


def _sorted_edges(ei):
    ei = ei.cpu().numpy()
    order = np.lexsort((ei[1], ei[0]))
    return ei[:, order]


def test_state_step_equals_step_batch(cuda_device):
    """Four consecutive edits per pair (incl. undoing an edit): rewards, solutions and graphs equal the Data-list path;
    the reward is scored on the ORIGINAL pair throughout."""
    g1, g2 = _pairs(cuda_device)
    env = GEDenv("ipfp", "AIDS-20-30", device=cuda_device)
    geds, ks, _ = env.solve_feasible_ged_batch(g1, g2, "ipfp")
    state = DeviceGEDState(env, g1, g2)
    assert torch.equal(state.solve(), torch.cat(geds))
    prev_l, prev_d = [float(g) for g in geds], torch.cat(geds)
    cur = list(g1)
    rng = np.random.default_rng(5)
    first = None
    for t in range(4):
        acts = np.stack([[rng.integers(0, g.num_nodes), 0] for g in cur]).T
        acts[1] = [(u + 1 + rng.integers(0, g.num_nodes - 1)) % g.num_nodes for u, g in zip(acts[0], cur)]
        if t == 0:
            first = acts.copy()
        if t == 3:
            acts = first[::-1].copy()                                  # undo the first edit, endpoints swapped
        a = torch.from_numpy(acts).to(cuda_device)
        rewards, cur, sols = env.step_batch(cur, g2, ks, a.cpu(), prev_l)
        r_dev, s_dev = state.step(a, prev_d)
        assert torch.equal(r_dev, torch.cat(rewards)) and torch.equal(s_dev, torch.cat(sols)), t
        prev_l, prev_d = [float(s) for s in sols], s_dev
        for want, got in zip(cur, state.edge_lists()):
            assert np.array_equal(_sorted_edges(want.edge_index), got.numpy()), t
    state.raise_on_error()
    b1 = state.batch1()
    ref = ppo_mod.DenseGraphBatch.from_data_list(cur, cuda_device)
    assert torch.equal(b1.adj, ref.adj) and torch.equal(b1.x, ref.x) and torch.equal(b1.mask, ref.mask)


def test_select_builds_successor_states(cuda_device):
    g1, g2 = _pairs(cuda_device, num=5, seed=902)
    env = GEDenv("ipfp", "AIDS-20-30", device=cuda_device)
    state = DeviceGEDState(env, g1, g2)
    parents = torch.tensor([4, 0, 0, 2, 4, 1])
    u = torch.tensor([0, 1, 2, 3, 4, 5])
    v = torch.tensor([7, 0, 9, 1, 2, 3])
    succ = state.select(parents, u, v)
    want_g1 = [env.toggle_edge(g1[int(p)], torch.tensor([int(a), int(b)])) for p, a, b in zip(parents, u, v)]
    for want, got in zip(want_g1, succ.edge_lists()):
        assert np.array_equal(_sorted_edges(want.edge_index), got.numpy())
    # solved on the successors, scored on the ORIGINAL parents' pairs
    ks = [env.construct_k(g1[int(p)], g2[int(p)]) for p in parents]
    want, _, _ = env.solve_feasible_ged_batch(want_g1, [g2[int(p)] for p in parents], "ipfp", ks)
    assert torch.equal(succ.solve(), torch.cat(want))
    assert torch.equal(state.solve(), torch.cat(env.solve_feasible_ged_batch(g1, g2, "ipfp")[0]))     # the parent state is untouched


def test_device_rollout_equals_list_rollout(cuda_device):
    """Same seed, same policy: the device-resident episode (no host round trips) and the Data-list episode sample the same
    actions, collect the same rewards and leave the same parameters after the PPO update."""
    g1, g2 = _pairs(cuda_device, num=8, seed=903)
    env = GEDenv("ipfp", "AIDS-20-30", device=cuda_device)
    geds, ks, _ = env.solve_feasible_ged_batch(g1, g2, "ipfp")
    tuples = [(a, b, k, float(g)) for a, b, k, g in zip(g1, g2, ks, geds)]
    torch.manual_seed(0)
    ppo_a = ppo_mod.PPO(ppo_mod.PPOConfig(update_timestep=4, k_epochs=2), cuda_device)
    ppo_b = ppo_mod.PPO(ppo_mod.PPOConfig(update_timestep=4, k_epochs=2), cuda_device)
    ppo_b.policy.load_state_dict(copy.deepcopy(ppo_a.policy.state_dict()))
    ppo_b.policy_old.load_state_dict(copy.deepcopy(ppo_a.policy_old.state_dict()))
    torch.manual_seed(11)
    _, total_a, loss_a = ppo_mod.rollout_episode(ppo_a, env, tuples, ppo_mod.Memory(), 4, 0, 4)
    state = DeviceGEDState(env, g1, g2)
    torch.manual_seed(11)
    _, total_b, loss_b = ppo_mod.rollout_episode_device(ppo_b, state, torch.cat(geds), ppo_mod.Memory(), 4, 0, 4)
    state.raise_on_error()
    assert np.array_equal(np.asarray(total_a, np.float32), total_b.cpu().numpy())
    assert len(loss_a) == len(loss_b) == 1 and abs(float(loss_a[0]) - float(loss_b[0])) <= 1e-6 * abs(float(loss_a[0]))
    for (name, pa), pb in zip(ppo_a.policy.named_parameters(), ppo_b.policy.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7), name
    assert torch.equal(state.adj1, DeviceGEDState(env, g1, g2).adj1)                                   # the start state is left untouched


def test_graph_captured_update_equals_eager_update(cuda_device, monkeypatch):
    """PPO.update with the K epochs replayed from a CUDA graph (captured after the first eager epoch; a second update of the
    same shapes reuses the graph with new inputs) leaves the parameters of the eager update."""
    g1, g2 = _pairs(cuda_device, num=8, seed=905)
    env = GEDenv("ipfp", "AIDS-20-30", device=cuda_device)
    geds, _, _ = env.solve_feasible_ged_batch(g1, g2, "ipfp")
    state = DeviceGEDState(env, g1, g2)
    results = []
    for graphs in ("1", "0"):
        monkeypatch.setenv("GED_B200_PPO_GRAPH", graphs)
        torch.manual_seed(0)
        ppo = ppo_mod.PPO(ppo_mod.PPOConfig(update_timestep=3, k_epochs=4), cuda_device)
        assert ppo.use_cuda_graph == (graphs == "1")
        torch.manual_seed(13)
        mem, t, losses = ppo_mod.Memory(), 0, []
        for _ in range(2):                                   # two updates: capture in the first, pure replay in the second
            t, _, loss = ppo_mod.rollout_episode_device(ppo, state, torch.cat(geds), mem, 3, t, 3)
            losses += loss
        assert len(ppo._graphs) == (1 if graphs == "1" else 0)
        results.append(([p.detach().clone() for p in ppo.policy.parameters()], [float(x) for x in losses]))
    (pa, la), (pb, lb) = results
    assert np.allclose(la, lb, rtol=1e-5)
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_device_beam_search_equals_list_beam_search(cuda_device):
    """The device-resident search and the Data-list search expand the same candidates and return the same best edit
    sequences (deterministic policy, both padded to the same node count)."""
    from ppo_bihyb_b200 import beam_search as bs
    from ppo_bihyb_b200.policy import ActorCritic
    g1, g2 = _pairs(cuda_device, num=4, seed=904)
    env = GEDenv("ipfp", "AIDS-20-30", device=cuda_device)
    geds, ks, _ = env.solve_feasible_ged_batch(g1, g2, "ipfp")
    tuples = [(a, b, k, float(g)) for a, b, k, g in zip(g1, g2, ks, geds)]
    torch.manual_seed(3)
    policy = ActorCritic(34, 64, False, 0, 3).to(cuda_device).eval()
    want = bs.beam_search_batch(policy, env, tuples, max_actions=3, beam_size=2)
    for search in (bs.beam_search_batch_device, bs.beam_search_batch_device_hostloop):     # sync-free search; round 1's host-loop search
        got = search(policy, env, tuples, max_actions=3, beam_size=2)
        for w, g in zip(want, got):
            assert g["reward"] == w["reward"] and g["solution"] == w["solution"], search.__name__
            if g["acts"] != w["acts"]:       # equal-reward alternatives may be ordered by probabilities that differ in the last bit
                assert len(g["acts"]) == len(w["acts"])
            else:
                assert np.allclose(np.asarray(g["probs"]), np.asarray(w["probs"]), rtol=1e-4)
    res = bs.evaluate(policy, env, [t + ({"ipfp": t[3]}, None) for t in tuples], max_steps=2, search_size=2)
    assert set(res) == {"reward", "ratio", "solution", "gap", "num_act", "time"} and "mean" in res["solution"]


def test_cli_trains_saves_and_evaluates_a_checkpoint(cuda_device, tmp_path, capsys):
    """The drivers end to end on synthetic pairs: two episodes of 8 pairs, a beam-search test after the second, a checkpoint
    under the reference's naming whose state_dict the evaluation driver (and the reference's ActorCritic) loads."""
    import glob
    from ppo_bihyb_b200 import cli
    common = ["--synthetic", "aids-20-30", "--solver_type", "ipfp", "--train_sample", "8", "--test_sample", "3", "--batch_size", "8",
              "--max_timesteps", "3", "--update_timestep", "3", "--k_epochs", "2", "--node_output_size", "64", "--gnn_layers", "3",
              "--search_size", "2", "--random_seed", "1"]
    cli.main(["train"] + common + ["--max_episodes", "2", "--log_interval", "1", "--test_interval", "2", "--save_dir", str(tmp_path),
                                   "--json_log"])
    out = capsys.readouterr().out
    assert "Episode 2" in out and "Evaluate on Test" in out and '"ratio/test"' in out
    saved = glob.glob(str(tmp_path / "PPO_ipfp_datasetaids-20-30_beam2_ratio*.pt"))
    if saved:                                   # written when the test ratio beats 0 (an untrained policy may not)
        sd = torch.load(saved[0], map_location="cpu")
        assert "state_encoder.siamese_gcn.conv_0.weight" in sd and sd["state_encoder.siamese_gcn.conv_0.weight"].shape == (34, 64)
        cli.main(["eval"] + common + ["--test_model_weight", saved[0]])
        assert '"solution"' in capsys.readouterr().out


def test_cli_evaluates_on_a_gedlib_directory(cuda_device, tmp_path, capsys):
    """The evaluation driver on a GEDLIB-style directory (synthetic .gxl files): dataset processing, tuple generation with
    reference-format cache files, device-resident beam search."""
    from _gxl import synthetic_raw_dir
    from ppo_bihyb_b200 import cli
    root = str(tmp_path / "gedlib")
    synthetic_raw_dir(root)
    cache = str(tmp_path / "cache")
    cli.main(["eval", "--dataset", "AIDS-10-16", "--dataset_root", root, "--cache_dir", cache, "--solver_type", "ipfp",
              "--train_sample", "4", "--test_sample", "3", "--max_timesteps", "2", "--search_size", "2",
              "--node_output_size", "64", "--gnn_layers", "3", "--random_seed", "0"])
    out = capsys.readouterr().out
    assert "ipfp ged mean=" in out and "BEAMSEARCH" in out and '"solution"' in out
    import os
    assert sorted(os.listdir(cache)) == sorted(f"{s}_{n}_AIDS-10-16_{r}.pt" for s in ("hungarian", "ipfp", "rrwm") for n, r in ((4, 0), (3, 1)))
