"""Row f-2: ``ppo_bihyb_b200.ppo.PPO.update`` against the UNMODIFIED reference ``PPO.update``
(ged_ppo_bihyb_train.py:158-221; golden made by oracle/gen_golden_ppo.py), and the data-parallel form (2 ranks, gloo, CPU)
against the single-process update on the full batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ppo_bihyb_b200 import ppo as ppo_mod
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.graphs import Data

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _small_state_dict():
    gold = np.load(os.path.join(GOLDEN, "policy.npz"))
    return {k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("sd/")}


def test_update_matches_reference_ppo():
    z = np.load(os.path.join(GOLDEN, "ppo.npz"))
    cfg = ppo_mod.PPOConfig(k_epochs=3, update_timestep=4, node_output_size=16)
    ppo = ppo_mod.PPO(cfg, torch.device("cpu"))
    sd = _small_state_dict()
    ppo.policy.load_state_dict(sd)
    ppo.policy_old.load_state_dict(sd)
    memory = ppo_mod.Memory()
    T = len(z["rewards"])
    d2 = Data(x=torch.from_numpy(z["x2"]), edge_index=torch.from_numpy(z["ei2"]))
    for t in range(T):
        g1 = Data(x=torch.from_numpy(z["x1"]), edge_index=torch.from_numpy(z[f"ei1_{t}"]))
        action = torch.from_numpy(z["acts"][t]).reshape(2, 1)
        with torch.no_grad():
            logp, _, _ = ppo.policy_old.evaluate([g1], [d2], action)
        memory.states.append(([g1], [d2]))
        memory.actions.append(action)
        memory.logprobs.append(logp)
        memory.rewards.append([torch.tensor([float(z["rewards"][t])])])
        memory.is_terminals.append([t == T - 1])
    got_old = torch.cat(memory.logprobs, dim=1).numpy()
    assert np.abs(got_old - z["old_logprobs"]).max() <= 1e-5 * np.abs(z["old_logprobs"]).max()      # fp32, 1e-5 relative
    closs = float(ppo.update(memory))
    assert abs(closs - float(z["critic_loss"])) <= 1e-4 * abs(float(z["critic_loss"]))
    # parameters after 3 Adam steps: Adam divides by sqrt(v) ~ |g|, so a 1e-7 relative difference in a gradient moves a
    # parameter by up to lr * O(1e-3); tolerance 2e-5 absolute (lr = 1e-3, encoder 1e-4)
    worst, moved = {}, 0
    for k, v in ppo.policy.state_dict().items():
        worst[k] = float((v - torch.from_numpy(z[f"after/{k}"])).abs().max())
        moved += int(float((torch.from_numpy(z[f"after/{k}"]) - sd[k]).abs().max()) > 1e-4)
    assert moved >= 15                                        # the golden update really trained the policy (21 tensors)
    # act1_resnet.first_linear.bias shifts all first-node scores by a constant: the soft-max ignores it, its true gradient is
    # zero and what Adam normalises is fp32 noise (~3e-8) -- on both sides.  It is excluded; everything else agrees.
    noise = {"actor_net.act1_resnet.first_linear.bias"}
    bad = {k: d for k, d in worst.items() if d > 2e-5 and k not in noise}
    assert not bad, bad
    for a, b in zip(ppo.policy.state_dict().values(), ppo.policy_old.state_dict().values()):
        assert torch.equal(a, b)                              # policy_old refreshed (:219)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _memory_for(ppo, pairs, seed):
    g = torch.Generator().manual_seed(seed)
    memory = ppo_mod.Memory()
    g1 = [Data.from_host(a) for a, _ in pairs]
    g2 = [Data.from_host(b) for _, b in pairs]
    for t in range(3):
        action = torch.stack([torch.tensor([int(torch.randint(0, a.n, (1,), generator=g)) for a, _ in pairs]),
                              torch.tensor([int(torch.randint(0, a.n, (1,), generator=g)) for a, _ in pairs])])
        action[1] = torch.where(action[1] == action[0], (action[1] + 1) % torch.tensor([a.n for a, _ in pairs]), action[1])
        with torch.no_grad():
            logp, _, _ = ppo.policy_old.evaluate(g1, g2, action)
        memory.states.append((g1, g2))
        memory.actions.append(action)
        memory.logprobs.append(logp)
        memory.rewards.append([torch.tensor([float(torch.randn(1, generator=g))]) for _ in pairs])
        memory.is_terminals.append([t == 2] * len(pairs))
    return memory


def _pairs():                                                 # equal sizes: the padded batch is the same on every rank layout
    rng = np.random.default_rng(3)
    return [(synthetic.molecule_like(24, rng), synthetic.molecule_like(22, rng)) for _ in range(4)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        cfg = ppo_mod.PPOConfig(k_epochs=2, node_output_size=16)
        ppo = ppo_mod.PPO(cfg, torch.device("cpu"))
        sd = _small_state_dict()
        ppo.policy.load_state_dict(sd)
        ppo.policy_old.load_state_dict(sd)
        full = _memory_for(ppo, _pairs(), seed=11)            # every rank builds the full memory, then keeps its shard
        mine = ppo_mod.Memory()
        sl = slice(rank * 2, rank * 2 + 2)
        for t in range(3):
            mine.states.append((full.states[t][0][sl], full.states[t][1][sl]))
            mine.actions.append(full.actions[t][:, sl])
            mine.logprobs.append(full.logprobs[t][:, sl])
            mine.rewards.append(full.rewards[t][sl])
            mine.is_terminals.append(full.is_terminals[t][sl])
        ppo.update(mine)
        q.put((rank, {k: v.numpy().copy() for k, v in ppo.policy.state_dict().items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_update_equals_single_process_update():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.set_num_threads(1)
    cfg = ppo_mod.PPOConfig(k_epochs=2, node_output_size=16)
    ppo = ppo_mod.PPO(cfg, torch.device("cpu"))
    sd = _small_state_dict()
    ppo.policy.load_state_dict(sd)
    ppo.policy_old.load_state_dict(sd)
    ppo.update(_memory_for(ppo, _pairs(), seed=11))
    for k, v in ppo.policy.state_dict().items():
        assert np.array_equal(got[0][k], got[1][k]), k                      # ranks stay in lock step
        if k == "actor_net.act1_resnet.first_linear.bias":                  # zero true gradient: Adam amplifies fp32 noise (see above)
            continue
        assert np.abs(got[0][k] - v.numpy()).max() <= 2e-5, k               # == the global-batch update (fp32 summation order)


import pytest


@pytest.mark.gpu
def test_rollout_episode_on_the_cuda_environment(cuda_device):
    """Three time steps for four pairs: one step_batch launch per time step, rewards consistent with the solutions the
    environment reports, one PPO update with a finite critic loss and changed parameters."""
    from ppo_bihyb_b200.ged_env import GEDenv
    torch.manual_seed(0)
    env = GEDenv("ipfp", "AIDS-20-30")
    ppo = ppo_mod.PPO(ppo_mod.PPOConfig(k_epochs=2, node_output_size=16, update_timestep=3), cuda_device)
    pairs = synthetic.make_pairs("aids-20-30", 4, seed_offset=801)
    g1 = [Data.from_host(a, cuda_device) for a, _ in pairs]
    g2 = [Data.from_host(b, cuda_device) for _, b in pairs]
    geds, ks, _ = env.solve_feasible_ged_batch(g1, g2, "ipfp")
    tuples = [(a, b, k, float(g)) for a, b, k, g in zip(g1, g2, ks, geds)]
    before = {k: v.clone() for k, v in ppo.policy.state_dict().items()}
    memory = ppo_mod.Memory()
    timestep, total, losses = ppo_mod.rollout_episode(ppo, env, tuples, memory, max_timesteps=3, timestep=0, update_timestep=3)
    assert timestep == 3 and len(losses) == 1 and bool(torch.isfinite(losses[0]))
    assert len(memory.states) == 0                                          # cleared after the update
    assert any(not torch.equal(before[k], v) for k, v in ppo.policy.state_dict().items())
    assert all(np.isfinite(t) for t in total)
