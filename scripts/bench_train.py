"""Row f-2 measurement (BASELINE.json configs[4]): PPO-BiHyb GED train steps -- rollout (policy act + one step_batch launch
per time step) and PPO update (K epochs, gradients all-reduced over the ranks) -- on synthetic AIDS-shaped pairs.

    python scripts/bench_train.py [--config aids-20-30] [--pairs 64] [--episodes 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py ...

Every rank rolls out `--pairs` pairs (weak scaling); prints environment steps/s over all ranks and the time split."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ppo_bihyb_b200 import ppo as ppo_mod
from ppo_bihyb_b200 import synthetic
from ppo_bihyb_b200.device_env import DeviceGEDState
from ppo_bihyb_b200.ged_env import GEDenv
from ppo_bihyb_b200.graphs import Data

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="aids-20-30")
ap.add_argument("--pairs", type=int, default=64)
ap.add_argument("--episodes", type=int, default=2)
ap.add_argument("--max-timesteps", type=int, default=10)
ap.add_argument("--host-state", action="store_true", help="roll out through GEDenv.step_batch on lists of Data graphs instead of the device-resident state")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                                   # identical initial policy on every rank
ppo = ppo_mod.PPO(ppo_mod.PPOConfig(update_timestep=args.max_timesteps), dev)
torch.manual_seed(1000 + rank)                         # independent action sampling
env = GEDenv("ipfp", "AIDS-20-30")
pairs = synthetic.make_pairs(args.config, args.pairs, seed_offset=700 + rank)
g1 = [Data.from_host(a, dev) for a, _ in pairs]
g2 = [Data.from_host(b, dev) for _, b in pairs]
geds, ks, _ = env.solve_feasible_ged_batch(g1, g2, "ipfp")
tuples = [(a, b, k, float(g)) for a, b, k, g in zip(g1, g2, ks, geds)]
state = None if args.host_state else DeviceGEDState(env, g1, g2)
greedy0 = torch.cat(geds)


def episode(timestep):
    if state is None:
        timestep, total, losses = ppo_mod.rollout_episode(ppo, env, tuples, memory, args.max_timesteps, timestep, args.max_timesteps)
        return timestep, sum(total) / len(total)
    timestep, total, losses = ppo_mod.rollout_episode_device(ppo, state, greedy0, memory, args.max_timesteps, timestep, args.max_timesteps)
    return timestep, total.mean()            # stays on the device until the end of the timed region

memory = ppo_mod.Memory()
timestep = 0
t_update = 0.0
orig_update = ppo.update


def timed_update(mem):
    # CUDA events, not synchronisation: the device-resident rollout never waits for the GPU inside an episode
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig_update(mem)
    e1.record()
    update_events.append((e0, e1))
    return out


update_events = []
ppo.update = timed_update
timestep, _ = episode(timestep)   # warm-up
torch.cuda.synchronize()
update_events.clear()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.time()
rewards = []
for _ in range(args.episodes):
    timestep, mean_reward = episode(timestep)
    rewards.append(mean_reward)
torch.cuda.synchronize()
wall = time.time() - t0
rewards = [float(r) for r in rewards]
if state is not None:
    state.raise_on_error()
t_update = sum(a.elapsed_time(b) for a, b in update_events) * 1e-3        # GPU time between the update's first and last kernel
dt = torch.tensor([wall, t_update], device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    flat = torch.cat([p.detach().reshape(-1) for p in ppo.policy.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    in_sync = bool(torch.equal(flat, ref))
else:
    in_sync = True
if rank == 0:
    steps = args.episodes * args.max_timesteps * args.pairs * world
    print(json.dumps({"metric": "ppo_bihyb_ged_env_steps_per_sec", "value": steps / float(dt[0]), "n_gpus": world, "config": args.config,
                      "state": "host lists" if args.host_state else "device resident", "pairs_per_gpu": args.pairs, "episodes": args.episodes, "max_timesteps": args.max_timesteps, "k_epochs": 10,
                      "seconds": float(dt[0]), "update_seconds": float(dt[1]), "rollout_seconds": float(dt[0] - dt[1]),
                      "mean_episode_reward": rewards, "parameters_identical_across_ranks": in_sync}))
if world > 1:
    dist.destroy_process_group()
