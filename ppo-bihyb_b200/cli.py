"""Command-line drivers around the path: training and beam-search evaluation of the bi-level GED policy.

    python -m ppo_bihyb_b200.cli train --config config/ppo_bihyb_ged.yaml [--dataset_root datasets/GEDLIB | --synthetic aids-20-30]
    python -m ppo_bihyb_b200.cli eval  --config config/ppo_bihyb_ged.yaml --test_model_weight pretrained/PPO_ipfp_....pt

Reference: the ``__main__`` blocks of ``ged_ppo_bihyb_train.py`` (:224-389, flags :392-449) and ``ged_ppo_bihyb_eval.py``
(:166-205).  Same flag names and defaults, same YAML semantics (every key of the file overrides the flag of that name; an
unknown key is an error), same log lines, same checkpoint naming (``PPO_{solver}_dataset{dataset}_beam{k}_ratio{r:.4f}.pt``,
a plain ``state_dict`` the reference loads unchanged).  What is different:

  * an episode rolls out ``batch_size`` pairs on the device (``ppo.rollout_episode_device``); the reference maps
    ``ged_env.step`` over a fork pool (:286-298);
  * under ``torchrun`` every rank trains on its own slice of the training tuples and the gradients are all-reduced;
  * ``--synthetic <config>`` trains / evaluates on the synthetic AIDS-shaped generator (no dataset in this image);
  * TensorBoard logging (TF-1.14 ``FileWriter``, :243-254) is not reproduced; the same scalars go to stdout as JSON lines
    with ``--json_log``.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

# flag -> (default, type) in the reference's order (ged_ppo_bihyb_train.py:397-431); booleans are store_true flags
FLAGS = {
    "solver_type": ("hungarian", str), "resource_limit": (600.0, float), "add_graph_features": (False, bool),
    "dataset": ("AIDS700nef", str), "gamma": (0.99, float), "train_sample": (50, int), "test_sample": (10, int),
    "search_size": (5, int), "learning_rate": (0.002, float), "lr_steps": ([], list), "batch_size": (1, int),
    "betas": ((0.9, 0.999), tuple), "max_episodes": (50000, int), "max_timesteps": (300, int), "update_timestep": (20, int),
    "k_epochs": (4, int), "eps_clip": (0.2, float), "one_hot_degree": (0, int), "batch_norm": (False, bool),
    "node_output_size": (16, int), "gnn_layers": (10, int), "config": (None, str), "random_seed": (None, int),
    "test_interval": (500, int), "log_interval": (100, int), "test_model_weight": ("", str),
}
EXTRA = {"dataset_root": (None, str), "synthetic": (None, str), "cache_dir": ("ged_data/cache", str), "json_log": (False, bool),
         "save_dir": (".", str)}


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser(description="GED bi-level PPO on the B200 solver: flags, or --config path/to/config.yaml")
    parser.add_argument("command", choices=("train", "eval"))
    for table in (FLAGS, EXTRA):
        for name, (default, typ) in table.items():
            if typ is bool:
                parser.add_argument("--" + name, action="store_true")
            elif typ in (list, tuple):
                parser.add_argument("--" + name, default=default, type=json.loads)
            else:
                parser.add_argument("--" + name, default=default, type=typ)
    args = parser.parse_args(argv)
    if args.config:
        import yaml
        path = args.config if os.path.exists(args.config) else os.path.join("config", args.config)      # the reference prefixes config/
        with open(path) as f:
            for key, val in (yaml.safe_load(f) or {}).items():
                assert hasattr(args, key), f"Unknown config key: {key}"
                setattr(args, key, val)
    return args


def _world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def build_environment(args, device):
    """Environment + (train tuples, test tuples) as ``generate_tuples`` returns them."""
    from . import synthetic
    from .ged_env import GEDenv
    from .graphs import Data
    if args.synthetic:
        env = GEDenv(args.solver_type, "AIDS-20-30", device=device)
        env.feature_dim = 34

        def tuples(num, seed):
            pairs = synthetic.make_pairs(args.synthetic, num, seed_offset=seed)
            g1 = [Data.from_host(a, device) for a, _ in pairs]
            g2 = [Data.from_host(b, device) for _, b in pairs]
            geds, ks, secs = env.solve_feasible_ged_batch(g1, g2, args.solver_type)
            return [(a, b, k, float(g), {args.solver_type: float(g)}, {args.solver_type: s}) for a, b, k, g, s in zip(g1, g2, ks, geds, secs)]
        return env, tuples(args.train_sample, 0), tuples(args.test_sample, 1)
    env = GEDenv(args.solver_type, args.dataset, device=device, dataset_root=args.dataset_root or "datasets/GEDLIB")
    train = env.generate_tuples(env.training_graphs, args.train_sample, 0, device, cache_dir=args.cache_dir)
    test = env.generate_tuples(env.val_graphs, args.test_sample, 1, device, cache_dir=args.cache_dir)
    return env, train, test


def _ppo_config(args):
    from .ppo import PPOConfig
    return PPOConfig(learning_rate=args.learning_rate, betas=tuple(args.betas), gamma=args.gamma, eps_clip=args.eps_clip,
                     k_epochs=args.k_epochs, lr_steps=list(args.lr_steps), update_timestep=args.update_timestep,
                     node_feature_dim=args.node_feature_dim, node_output_size=args.node_output_size, batch_norm=args.batch_norm,
                     one_hot_degree=args.one_hot_degree, gnn_layers=args.gnn_layers)


def train(args):
    from . import ppo as ppo_mod
    from .beam_search import evaluate
    from .device_env import DeviceGEDState
    world, rank, local = _world()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    if args.random_seed is not None:
        random.seed(args.random_seed)
        np.random.seed(args.random_seed)
    torch.manual_seed(args.random_seed if args.random_seed is not None else 0)        # identical initial policy on every rank
    env, tuples_train, tuples_test = build_environment(args, device)
    args.node_feature_dim = env.feature_dim
    trainer = ppo_mod.PPO(_ppo_config(args), device)
    base_seed = args.random_seed if args.random_seed is not None else 0
    torch.manual_seed(base_seed + 1000 * (rank + 1))                                   # independent action sampling per rank
    mine = tuples_train[rank::world] if world > 1 else tuples_train                   # every rank trains on its own slice
    if len(mine) == 0:
        raise ValueError(f'rank {rank} of {world} has no training tuples: train_sample={len(tuples_train)} < world size')
    memory = ppo_mod.Memory()
    best_test_ratio, running_reward, critic_loss, timestep = 0.0, 0.0, [], 0
    prev_time = time.time()
    states = {}
    for i_episode in range(1, args.max_episodes + 1):
        first = ((i_episode - 1) * args.batch_size) % len(mine)                        # the reference's round-robin over the tuples (:273)
        idx = tuple((first + b) % len(mine) for b in range(args.batch_size))
        if idx not in states:
            batch = [mine[i] for i in idx]
            states[idx] = (DeviceGEDState(env, [t[0] for t in batch], [t[1] for t in batch]),
                           torch.tensor([float(t[3]) for t in batch], device=device))
        state, greedy = states[idx]
        timestep, total, losses = ppo_mod.rollout_episode_device(trainer, state, greedy, memory, args.max_timesteps, timestep,
                                                                 args.update_timestep)
        running_reward += float(total.mean())
        critic_loss += losses
        if i_episode % args.log_interval == 0 and rank == 0:
            closs = float(torch.stack(critic_loss).mean()) if critic_loss else -1.0
            avg_time = (time.time() - prev_time) / args.log_interval
            prev_time = time.time()
            print(f"Episode {i_episode} \t avg length: {float(args.max_timesteps):.2f} \t critic mse: {closs:.4f} \t "
                  f"reward: {running_reward / args.log_interval:.4f} \t time per episode: {avg_time:.2f}", flush=True)
            if args.json_log:
                print(json.dumps({"episode": i_episode, "timestep": timestep, "reward/train": running_reward / args.log_interval,
                                  "critic mse/train": closs, "time/train": avg_time,
                                  "lr": [g["lr"] for g in trainer.optimizer.param_groups]}), flush=True)
            running_reward, critic_loss = 0.0, []
        elif i_episode % args.log_interval == 0:
            running_reward, critic_loss = 0.0, []
        if i_episode % args.test_interval == 0:
            state.raise_on_error()
            if rank == 0:
                t_test = time.time()
                print("########## Evaluate on Test ##########")
                test_dict = evaluate(trainer.policy, env, tuples_test, args.max_timesteps, args.search_size, verbose=True)
                print("########## Evaluate complete ##########", flush=True)
                if args.json_log:
                    print(json.dumps({"episode": i_episode, **{f"{k}/test": v["mean"] for k, v in test_dict.items()}}), flush=True)
                prev_time += time.time() - t_test
                if test_dict["ratio"]["mean"] > best_test_ratio:
                    best_test_ratio = test_dict["ratio"]["mean"]
                    name = f"PPO_{args.solver_type}_dataset{args.synthetic or args.dataset}_beam{args.search_size}_ratio{best_test_ratio:.4f}.pt"
                    torch.save(trainer.policy.state_dict(), os.path.join(args.save_dir, name))
            if world > 1:
                dist.barrier()
    if world > 1:
        dist.destroy_process_group()
    return best_test_ratio


def evaluate_checkpoint(args):
    from .beam_search import evaluate
    from .policy import ActorCritic
    device = torch.device("cuda", _world()[2])
    torch.cuda.set_device(device)
    if args.random_seed is not None:
        random.seed(args.random_seed)
        np.random.seed(args.random_seed)
        torch.manual_seed(args.random_seed)
    env, _, tuples_test = build_environment(args, device)
    policy = ActorCritic(env.feature_dim, args.node_output_size, args.batch_norm, args.one_hot_degree, args.gnn_layers).to(device)
    if args.test_model_weight:
        policy.load_state_dict(torch.load(args.test_model_weight, map_location=device))
    with torch.no_grad():
        return evaluate(policy.eval(), env, tuples_test, args.max_timesteps, args.search_size, verbose=True)


def main(argv=None):
    args = parse_arguments(argv)
    if args.command == "train":
        train(args)
    else:
        res = evaluate_checkpoint(args)
        print(json.dumps({k: v["mean"] for k, v in res.items()}))


if __name__ == "__main__":
    main(sys.argv[1:])
