"""PPO for the bi-level GED policy with data-parallel rollouts (BASELINE.json configs[4]).

Reference: ``ged_ppo_bihyb_train.py`` -- ``Memory`` :83-98, ``PPO`` :129-221 (clipped surrogate, K epochs, Adam with the
encoder at lr/10), rollout loop :270-321.  What is different:

  * the environment step of a whole batch of pairs is ONE ``GEDenv.step_batch`` launch (the reference maps
    ``ged_env.step`` over a fork pool, :286-298);
  * with ``torch.distributed`` initialised, every rank rolls out its own shard of the training pairs; ``update`` all-gathers
    the discounted rewards' first two moments (so the normalisation is the global one) and all-reduces the gradients as
    one NCCL call per epoch on a flat gradient bucket the parameters' ``.grad`` are views of -- the only inter-GPU traffic
    of a train step.
"""
import os
from dataclasses import dataclass, field
from typing import List, Sequence

import torch
import torch.distributed as dist
from torch import nn

from .policy import ActorCritic, DenseGraphBatch


class Memory:
    def __init__(self):
        self.actions, self.states, self.candidates, self.logprobs, self.rewards, self.is_terminals = [], [], [], [], [], []

    def clear_memory(self):
        for lst in (self.actions, self.states, self.candidates, self.logprobs, self.rewards, self.is_terminals):
            del lst[:]


@dataclass
class PPOConfig:
    """The keys of ``config/ppo_bihyb_ged.yaml`` that PPO reads (same names and defaults)."""
    learning_rate: float = 0.001
    betas: tuple = (0.9, 0.999)
    gamma: float = 0.95
    eps_clip: float = 0.1
    k_epochs: int = 10
    lr_steps: list = field(default_factory=list)
    update_timestep: int = 10
    node_feature_dim: int = 34
    node_output_size: int = 64
    batch_norm: bool = False
    one_hot_degree: int = 0
    gnn_layers: int = 3


def _distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class _EpochGraph:
    """One PPO epoch's forward + loss + backward as a CUDA graph over static copies of the update's inputs.  The K epochs of an
    update evaluate the same states with the same shapes (and so do the following updates of a training run), while the eager
    epoch is ~1500 kernel launches of a few microseconds each: launch-bound.  Captured after one eager epoch of the same shapes;
    the optimiser step and the gradient all-reduce stay outside (eager), so the graph holds no collective."""

    def __init__(self, ppo, inputs):
        self.static = [t.clone() for t in inputs]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.critic_loss = ppo._epoch_forward_backward(*self._unflatten(self.static))

    @staticmethod
    def flatten(g1, g2, actions, logprobs, rewards):
        return [g1.x, g1.adj, g1.mask, g1.num_nodes, g2.x, g2.adj, g2.mask, g2.num_nodes, actions, logprobs, rewards]

    @staticmethod
    def _unflatten(t):
        return DenseGraphBatch(*t[0:4]), DenseGraphBatch(*t[4:8]), t[8], t[9], t[10]

    @staticmethod
    def signature(inputs):
        return tuple((tuple(t.shape), t.dtype) for t in inputs)

    def load(self, inputs):
        for dst, src in zip(self.static, inputs):
            dst.copy_(src)

    def replay(self):
        self.graph.replay()
        return self.critic_loss


class PPO:
    def __init__(self, args: PPOConfig, device):
        self.lr, self.betas, self.gamma, self.eps_clip, self.K_epochs = args.learning_rate, args.betas, args.gamma, args.eps_clip, args.k_epochs
        self.device = device
        ac_params = args.node_feature_dim, args.node_output_size, args.batch_norm, args.one_hot_degree, args.gnn_layers
        self.policy = ActorCritic(*ac_params).to(device)
        self.optimizer = torch.optim.Adam(
            [{"params": self.policy.actor_net.parameters()},
             {"params": self.policy.value_net.parameters()},
             {"params": self.policy.state_encoder.parameters(), "lr": self.lr / 10}],
            lr=self.lr, betas=self.betas)
        if len(args.lr_steps) > 0:
            self.lr_scheduler = torch.optim.lr_scheduler.MultiStepLR(
                self.optimizer, [s // args.update_timestep for s in args.lr_steps], gamma=0.1)
        else:
            self.lr_scheduler = None
        self.policy_old = ActorCritic(*ac_params).to(device)
        self.policy_old.load_state_dict(self.policy.state_dict())
        self.MseLoss = nn.MSELoss()
        # One flat gradient bucket, every p.grad a view into it: zeroing is one fill, the data-parallel reduction is one
        # NCCL call on the bucket itself (no gather / scatter copies of the 77 parameter gradients per epoch).
        self.use_cuda_graph = os.environ.get("GED_B200_PPO_GRAPH", "1") != "0"
        self._graphs = {}
        params = list(self.policy.parameters())
        self._flat_grad = torch.zeros(sum(p.numel() for p in params), device=device)
        o = 0
        for p in params:
            p.grad = self._flat_grad[o:o + p.numel()].view_as(p)
            o += p.numel()

    def _epoch_forward_backward(self, old_graph_1, old_graph_2, old_actions, old_logprobs, rewards):
        """Forward, clipped-surrogate + critic + entropy loss and backward of one epoch (ged_ppo_bihyb_train.py:192-214); the
        gradients land in the flat bucket.  Returns the epoch's critic loss (detached)."""
        logprobs, state_values, dist_entropy = self.policy.evaluate(old_graph_1, old_graph_2, old_actions)
        ratios = torch.exp(logprobs - old_logprobs)
        advantages = rewards - state_values.detach()
        surr1 = ratios * advantages
        surr2 = torch.clamp(ratios, 1 - self.eps_clip, 1 + self.eps_clip) * advantages
        actor_loss = -torch.min(surr1, surr2)
        critic_loss = self.MseLoss(state_values, rewards)
        entropy_reg = -0.01 * dist_entropy
        self._flat_grad.zero_()
        (actor_loss + critic_loss + entropy_reg).mean().backward()      # accumulates in place into the bucket views
        return critic_loss.detach().mean()

    def _epoch_graph(self, inputs, capture: bool = False):
        """The captured epoch for these input shapes, loaded with these inputs; None when there is none (yet)."""
        if not self.use_cuda_graph or inputs[0].device.type != "cuda":
            return None
        sig = _EpochGraph.signature(inputs)
        g = self._graphs.get(sig)
        if g is None:
            if not capture:
                return None
            if len(self._graphs) >= 4:               # a handful of shapes per run (the padded node count of the training batch)
                self._graphs.pop(next(iter(self._graphs)))
            g = self._graphs[sig] = _EpochGraph(self, inputs)
            return g
        g.load(inputs)
        return g

    def _discounted_rewards(self, memory):
        with torch.no_grad():
            _, state_values, _ = self.policy.evaluate(*memory.states[-1], memory.actions[-1].to(self.device))
        discounted = state_values
        rewards = []
        for reward, is_terminal in zip(reversed(memory.rewards), reversed(memory.is_terminals)):
            if torch.is_tensor(reward):          # device-resident rollout: one [B] tensor per time step
                reward = reward.to(device=self.device, dtype=torch.float32)
            else:
                reward = torch.as_tensor([float(r) for r in reward], dtype=torch.float32, device=self.device)
            term = torch.as_tensor([float(t) for t in is_terminal], dtype=torch.float32, device=self.device)
            discounted = reward + self.gamma * (discounted * (1 - term))
            rewards.insert(0, discounted)
        return torch.cat(rewards)

    def _normalise(self, rewards):
        """(r - mean) / (std + 1e-5) with torch's unbiased std, over the rewards of ALL ranks."""
        if not _distributed():
            return (rewards - rewards.mean()) / (rewards.std() + 1e-5)
        stats = torch.stack((rewards.sum(), (rewards * rewards).sum(), torch.tensor(float(rewards.numel()), device=rewards.device))).double()
        dist.all_reduce(stats)
        n = stats[2]
        mean = stats[0] / n
        var = (stats[1] - n * mean * mean) / (n - 1)
        return ((rewards.double() - mean) / (var.clamp(min=0).sqrt() + 1e-5)).float()

    def update(self, memory) -> torch.Tensor:
        """``PPO.update`` (ged_ppo_bihyb_train.py:158-221).  Returns the mean critic loss of the K epochs."""
        rewards = self._normalise(self._discounted_rewards(memory))
        if _distributed():              # the gradient mean below is the global-batch gradient only for equal shards
            count = torch.tensor([float(rewards.numel())], device=self.device)
            lo, hi = count.clone(), count.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if float(lo) != float(hi):
                raise ValueError(f'PPO.update: ranks hold different numbers of transitions ({int(lo)}..{int(hi)})')
        if isinstance(memory.states[0][0], DenseGraphBatch):     # device-resident rollout: snapshots of the dense batches
            old_graph_1 = DenseGraphBatch.cat([state[0] for state in memory.states])
            old_graph_2 = DenseGraphBatch.cat([state[1] for state in memory.states])
        else:
            old_graph_1, old_graph_2 = [], []
            for state in memory.states:
                old_graph_1 += state[0]
                old_graph_2 += state[1]
            old_graph_1 = DenseGraphBatch.from_data_list(old_graph_1, self.device)      # padded once, evaluated K times
            old_graph_2 = DenseGraphBatch.from_data_list(old_graph_2, self.device)
        old_actions = torch.cat(memory.actions, dim=1).to(self.device)
        old_logprobs = torch.cat(memory.logprobs, dim=1).to(self.device).detach()
        critic_loss_sum = 0
        inputs = _EpochGraph.flatten(old_graph_1, old_graph_2, old_actions, old_logprobs, rewards)
        graph = self._epoch_graph(inputs)            # None: eager (CPU, graphs switched off, or the first epoch of a new shape)
        for epoch in range(self.K_epochs):
            if graph is None and epoch == 1:         # one eager epoch of these shapes has run: capture the rest
                graph = self._epoch_graph(inputs, capture=True)
            if graph is not None:
                critic_loss = graph.replay()
            else:
                critic_loss = self._epoch_forward_backward(old_graph_1, old_graph_2, old_actions, old_logprobs, rewards)
            critic_loss_sum = critic_loss_sum + critic_loss
            if _distributed():          # equal shard sizes: the mean of the rank gradients is the global-batch gradient
                dist.all_reduce(self._flat_grad)
                self._flat_grad /= dist.get_world_size()
            self.optimizer.step()
        if self.lr_scheduler:
            self.lr_scheduler.step()
        self.policy_old.load_state_dict(self.policy.state_dict())
        return critic_loss_sum / self.K_epochs


def rollout_episode(ppo: PPO, env, tuples: Sequence, memory: Memory, max_timesteps: int, timestep: int, update_timestep: int):
    """One episode of the reference's training loop (:270-321) for a batch of pairs: ``max_timesteps`` edits per pair,
    every environment step of the batch in one ``step_batch`` launch, ``ppo.update`` whenever ``timestep`` hits the
    update interval.  ``tuples[b] = (graph_1, graph_2, ori_k, ori_greedy, ...)``.
    Returns (new timestep, summed reward per pair, list of critic losses)."""
    g1 = [t[0] for t in tuples]
    g2 = [t[1] for t in tuples]
    ks = [t[2] for t in tuples]
    greedy = [t[3] for t in tuples]
    total = [0.0] * len(tuples)
    losses = []
    for t in range(max_timesteps):
        timestep += 1
        with torch.no_grad():
            action_batch = ppo.policy_old.act(g1, g2, memory)
        rewards, g1, greedy = env.step_batch(g1, g2, ks, action_batch.cpu(), greedy)
        memory.rewards.append(rewards)
        memory.is_terminals.append([t == max_timesteps - 1] * len(tuples))
        total = [a + float(r) for a, r in zip(total, rewards)]
        if timestep % update_timestep == 0:
            losses.append(ppo.update(memory))
            memory.clear_memory()
    return timestep, total, losses


def rollout_episode_device(ppo: PPO, state, greedy: torch.Tensor, memory: Memory, max_timesteps: int, timestep: int,
                           update_timestep: int):
    """:func:`rollout_episode` on a :class:`device_env.DeviceGEDState`: the graphs, the actions, the rewards and the PPO memory
    stay on the device and nothing synchronises -- the policy forward reads the state's dense batches, ``state.step`` is one
    solver call plus the in-place edge toggles.  ``state`` is the episode's start (left untouched); ``greedy`` [B] are the
    solutions of the unedited pairs.  Returns (new timestep, summed reward per pair [B] on the device, critic losses)."""
    st = state.clone()
    greedy = greedy.to(st.device, torch.float32)
    total = torch.zeros(len(st), device=st.device)
    losses = []
    for t in range(max_timesteps):
        timestep += 1
        with torch.no_grad():
            actions = ppo.policy_old.act(st.batch1(), st.batch2(), memory)
        rewards, greedy = st.step(actions, greedy)
        memory.rewards.append(rewards)
        memory.is_terminals.append([t == max_timesteps - 1] * len(st))
        total += rewards
        if timestep % update_timestep == 0:
            losses.append(ppo.update(memory))
            memory.clear_memory()
    state.status |= st.status
    return timestep, total, losses
