"""Upper-level policy of PPO-BiHyb for GED, as dense batched PyTorch (the caller that feeds
the lower-level solver).  No torch_geometric / torch_scatter: every graph of a batch is padded to the batch's largest
node count and the GCN runs as a normalised-adjacency matmul, so all beam states of ALL test pairs go through one forward.

Reference (file:line in /root/reference):
  ``GraphEncoder``            src/ged_ppo_bihyb_model.py:11-60      ``ActorNet``   :63-119     ``CriticNet``  :122-145
  ``GCN`` / ``ResNetBlock``   src/pyg_graph_models.py:49-86, 8-46   ``GraphAttentionPooling`` :89-128
  ``TensorNetworkModule``     src/pyg_graph_models.py:174-222       ``Sinkhorn.forward_log``   utils/sinkhorn.py:143-239
  ``ActorCritic``             ged_ppo_bihyb_train.py:101-127

Module / parameter names equal the reference's, so its checkpoints (``pretrained/PPO_ipfp_dataset*.pt``) load with
``load_state_dict`` unchanged (GCN weights are stored (in, out), the PyG-1.7 layout).

Differences on purpose: padded node slots are masked out of the action soft-max (the reference never batches graphs of
different sizes in one call -- all beam states of a pair have the pair's node count -- so per pair the probabilities are
identical).
"""
from typing import List, Sequence, Tuple

import torch
from torch import nn
from torch.distributions import Categorical


class DenseGraphBatch:
    """B graphs padded to n = max node count: ``x`` [B, n, F], ``adj`` [B, n, n] (edge multiplicities, as PyG sums
    duplicate edges), ``mask`` [B, n] bool, ``num_nodes`` [B]."""

    def __init__(self, x, adj, mask, num_nodes):
        self.x, self.adj, self.mask, self.num_nodes = x, adj, mask, num_nodes

    @staticmethod
    def cat(batches: Sequence["DenseGraphBatch"]) -> "DenseGraphBatch":
        """Concatenate along the batch dimension (all padded to the same node count: snapshots of one device state)."""
        return DenseGraphBatch(torch.cat([b.x for b in batches]), torch.cat([b.adj for b in batches]),
                               torch.cat([b.mask for b in batches]), torch.cat([b.num_nodes for b in batches]))

    @staticmethod
    def from_data_list(graphs: Sequence, device) -> "DenseGraphBatch":
        B = len(graphs)
        ns = [int(g.x.shape[0]) for g in graphs]
        n = max(ns)
        F = int(graphs[0].x.shape[1])
        x = torch.zeros(B, n, F, device=device)
        bidx, rows, cols = [], [], []
        for b, g in enumerate(graphs):
            x[b, :ns[b]] = g.x.to(device=device, dtype=torch.float32)
            ei = g.edge_index
            bidx.append(torch.full((ei.shape[1],), b, dtype=torch.long))
            rows.append(ei[0].to("cpu"))
            cols.append(ei[1].to("cpu"))
        adj = torch.zeros(B, n, n, device=device)
        if bidx:
            idx = (torch.cat(bidx).to(device), torch.cat(rows).to(device), torch.cat(cols).to(device))
            adj.index_put_(idx, torch.ones(idx[0].shape[0], device=device), accumulate=True)
        num_nodes = torch.tensor(ns, device=device)
        mask = torch.arange(n, device=device).unsqueeze(0) < num_nodes.unsqueeze(1)
        return DenseGraphBatch(x, adj, mask, num_nodes)


class DenseGCNConv(nn.Module):
    """``torch_geometric.nn.GCNConv`` (PyG 1.7) on a padded dense batch: out = D^-1/2 (A + I)^T D^-1/2 X W + b with one
    self loop per node (existing self loops are replaced by it) and D = in-degree incl. the self loop."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x, adj):
        n = adj.shape[-1]
        eye = torch.eye(n, device=adj.device, dtype=adj.dtype)
        a_hat = adj * (1 - eye) + eye                       # [B, r, c]
        deg = a_hat.sum(dim=1)                              # in-degree of c
        dinv = deg.pow(-0.5)
        xw = torch.matmul(x, self.weight)
        out = torch.bmm(a_hat.transpose(1, 2) * dinv.unsqueeze(2) * dinv.unsqueeze(1), xw)
        return out + self.bias


class GCN(nn.Module):
    def __init__(self, num_in_feats, num_out_feats, num_layers=3, batch_norm=True):
        super().__init__()
        self.num_layers = num_layers
        for l in range(num_layers):
            self.add_module(f"conv_{l}", DenseGCNConv(num_in_feats if l == 0 else num_out_feats, num_out_feats))
            self.add_module(f"norm_{l}", nn.BatchNorm1d(num_out_feats) if batch_norm else nn.Identity())

    def forward(self, batch: DenseGraphBatch):
        x = batch.x
        for l in range(self.num_layers):
            x = getattr(self, f"conv_{l}")(x, batch.adj)
            norm = getattr(self, f"norm_{l}")
            if isinstance(norm, nn.BatchNorm1d):            # statistics over the real nodes only, as on PyG's flat node list
                flat = norm(x[batch.mask])
                x = torch.zeros_like(x).masked_scatter(batch.mask.unsqueeze(-1), flat)
            x = torch.relu(x)
        return x * batch.mask.unsqueeze(-1)                 # to_dense_batch pads with zeros


class GraphAttentionPooling(nn.Module):
    def __init__(self, feat_dim):
        super().__init__()
        self.weight_matrix = nn.Parameter(torch.empty(feat_dim, feat_dim))
        nn.init.xavier_uniform_(self.weight_matrix)

    def forward(self, x, mask, num_nodes):
        mean = x.sum(dim=1) / num_nodes.unsqueeze(-1).to(x.dtype)
        transformed_global = torch.tanh(torch.mm(mean, self.weight_matrix))
        coefs = torch.sigmoid((x * transformed_global.unsqueeze(1) * 10).sum(dim=-1)) * mask
        return (coefs.unsqueeze(-1) * x).sum(dim=1)


def log_sinkhorn(s, nrows, ncols, max_iter=20, tau=0.005):
    """``Sinkhorn.forward_log`` (utils/sinkhorn.py:143-239) with per-sample sizes, batched: entries outside a sample's
    nrows x ncols block stay at -inf (they contribute nothing to any logsumexp), iterations alternate row / column
    normalisation starting with rows.  Without autograd on a CUDA device this is ``ged_sinkhorn_batch``."""
    B, n1, n2 = s.shape
    if s.is_cuda and not torch.is_grad_enabled():
        # inference (act / propose): the fused kernel -- all passes in shared memory, one launch instead of ~4 per pass
        from . import solver
        return solver.sinkhorn_batch(s, nrows, ncols, torch.full((B,), float(tau), device=s.device), max_iter)
    valid = (torch.arange(n1, device=s.device)[None, :, None] < nrows[:, None, None]) & \
            (torch.arange(n2, device=s.device)[None, None, :] < ncols[:, None, None])
    neg_inf = torch.full_like(s, -float("inf"))
    log_s = torch.where(valid, s / tau, neg_inf)
    for i in range(max_iter):
        log_sum = torch.logsumexp(log_s, 2 if i % 2 == 0 else 1, keepdim=True)
        log_s = torch.where(valid, log_s - log_sum, neg_inf)
    return torch.exp(log_s)


class GraphEncoder(nn.Module):
    def __init__(self, node_feature_dim, node_output_size, batch_norm, one_hot_degree, num_layers=10):
        super().__init__()
        if one_hot_degree:
            raise NotImplementedError("one_hot_degree > 0: the GED configs use 0 (degree is already in the node features)")
        self.siamese_gcn = GCN(node_feature_dim, node_output_size, num_layers=num_layers, batch_norm=batch_norm)
        self.att = GraphAttentionPooling(node_output_size)
        self.sinkhorn_iters, self.sinkhorn_tau = 20, 0.005

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, input_graphs_1, input_graphs_2):
        """Lists of graphs, or two pre-built :class:`DenseGraphBatch` (PPO.update evaluates the same states K times)."""
        b1 = input_graphs_1 if isinstance(input_graphs_1, DenseGraphBatch) else DenseGraphBatch.from_data_list(input_graphs_1, self.device)
        b2 = input_graphs_2 if isinstance(input_graphs_2, DenseGraphBatch) else DenseGraphBatch.from_data_list(input_graphs_2, self.device)
        node_feat_1, node_feat_2 = self.siamese_gcn(b1), self.siamese_gcn(b2)
        sim_mat = torch.bmm(node_feat_1, node_feat_2.transpose(1, 2))
        sim_mat = log_sinkhorn(sim_mat, b1.num_nodes, b2.num_nodes, self.sinkhorn_iters, self.sinkhorn_tau)
        diff_feat = node_feat_1 - torch.bmm(sim_mat, node_feat_2)
        global_feat_1 = self.att(node_feat_1, b1.mask, b1.num_nodes)
        global_feat_2 = self.att(node_feat_2, b2.mask, b2.num_nodes)
        self.last_mask = b1.mask
        return diff_feat, global_feat_1, global_feat_2


class ResNetBlock(nn.Module):
    """``ResNetBlock`` of the reference without batch norm; ``last_linear`` exists in the checkpoints but the reference's
    forward never applies it (src/pyg_graph_models.py:43-46) -- kept as an unused parameter for state-dict compatibility."""

    def __init__(self, num_in_feats, num_out_feats, num_layers=3, batch_norm=False):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("batch_norm in the actor head is not used by the GED checkpoints")
        self.first_linear = nn.Linear(num_in_feats, num_out_feats)
        self.last_linear = nn.Linear(num_out_feats, num_out_feats)
        seq = [nn.ReLU()]
        for _ in range(1, num_layers - 1):
            seq += [nn.Linear(num_out_feats, num_out_feats), nn.ReLU()]
        self.sequential = nn.Sequential(*seq)

    def forward(self, inp):
        x1 = self.first_linear(inp)
        return self.sequential(x1) + x1


class ActorNet(nn.Module):
    def __init__(self, state_feature_size, batch_norm):
        super().__init__()
        self.act1_resnet = ResNetBlock(state_feature_size, 1, batch_norm=batch_norm)
        self.act2_query = nn.Linear(state_feature_size, state_feature_size, bias=False)

    def select_nodes(self, state_feat, node_mask, greedy_sel_num, prev_act=None):
        """``_select_node`` with ``greedy_sel_num > 0`` (src/ged_ppo_bihyb_model.py:92-112): top-k node ids and their
        probabilities.  ``node_mask`` [B, n] marks the real nodes."""
        if prev_act is None:
            act_scores = self.act1_resnet(state_feat).squeeze(-1)
            mask = torch.zeros_like(act_scores)
        else:
            rows = torch.arange(len(prev_act), device=state_feat.device)
            act_query = torch.tanh(self.act2_query(state_feat[rows, prev_act, :]))
            act_scores = (act_query.unsqueeze(1) * state_feat).sum(dim=-1)
            # scatter_ takes the scalar as a kernel argument (an indexed assignment of a Python float copies it from pageable
            # host memory: not capturable in a CUDA graph)
            mask = torch.zeros_like(act_scores).scatter_(1, prev_act.unsqueeze(1), -float("inf"))
        mask = mask.masked_fill(~node_mask, -float("inf"))
        act_probs = torch.softmax(act_scores + mask, dim=1)
        if greedy_sel_num > 0:
            acts = torch.argsort(act_probs, dim=-1, descending=True)[:, :greedy_sel_num]
            return acts, torch.gather(act_probs, 1, acts)
        return act_probs

    def forward(self, state_feat, node_mask, known_action=None):
        """``_act`` (src/ged_ppo_bihyb_model.py:84-90): sample (or score the known) first node, then the second node given
        the first.  Returns actions [2, B], log-probabilities [2, B], entropy [B]."""
        if known_action is None:
            known_action = (None, None)
        out = []
        prev = None
        for step in range(2):
            # validate_args=False: the constraint check reads a flag back to the host (a synchronisation per call, and not
            # capturable in a CUDA graph); the probabilities are a masked soft-max
            dist = Categorical(probs=self.select_nodes(state_feat, node_mask, 0, prev_act=prev), validate_args=False)
            act = dist.sample() if known_action[step] is None else known_action[step]
            out.append((act, dist.log_prob(act), dist.entropy()))
            prev = act
        return torch.stack((out[0][0], out[1][0])), torch.stack((out[0][1], out[1][1])), out[0][2] + out[1][2]


class TensorNetworkModule(nn.Module):
    def __init__(self, input_features, tensor_neurons):
        super().__init__()
        self.input_features, self.tensor_neurons = input_features, tensor_neurons
        self.weight_matrix = nn.Parameter(torch.empty(input_features, input_features, tensor_neurons))
        self.weight_matrix_block = nn.Parameter(torch.empty(tensor_neurons, 2 * input_features))
        self.bias = nn.Parameter(torch.empty(tensor_neurons, 1))
        for p in (self.weight_matrix, self.weight_matrix_block, self.bias):
            nn.init.xavier_uniform_(p)

    def forward(self, e1, e2):
        B, F = e1.shape[0], self.input_features
        scoring = torch.matmul(e1, self.weight_matrix.view(F, -1)).view(B, F, -1).permute(0, 2, 1)
        scoring = torch.matmul(scoring, e2.view(B, F, 1)).view(B, -1)
        block = torch.mm(self.weight_matrix_block, torch.cat((e1, e2), 1).t()).t()
        return torch.relu(scoring + block + self.bias.view(-1))


class CriticNet(nn.Module):
    def __init__(self, state_feature_size, batch_norm):
        super().__init__()
        self.tensor_net = TensorNetworkModule(state_feature_size, state_feature_size)
        self.scoring_layer = nn.Sequential(nn.Linear(state_feature_size, 16), nn.ReLU(), nn.Linear(16, 1), nn.Tanh())

    def forward(self, global_feat_1, global_feat_2):
        return self.scoring_layer(self.tensor_net(global_feat_1, global_feat_2)).squeeze(-1)


class ActorCritic(nn.Module):
    """Same composition and parameter names as ``ged_ppo_bihyb_train.py:101-108``."""

    def __init__(self, node_feature_dim, node_output_size, batch_norm, one_hot_degree, gnn_layers):
        super().__init__()
        self.state_encoder = GraphEncoder(node_feature_dim, node_output_size, batch_norm, one_hot_degree, gnn_layers)
        self.actor_net = ActorNet(node_output_size, batch_norm)
        self.value_net = CriticNet(node_output_size, batch_norm)

    def act(self, inp_graph_1: List, inp_graph_2: List, memory):
        """``ActorCritic.act`` (ged_ppo_bihyb_train.py:112-120): sample one edge edit per state and record it."""
        diff_feat, _, _ = self.state_encoder(inp_graph_1, inp_graph_2)
        actions, action_logits, _ = self.actor_net(diff_feat, self.state_encoder.last_mask)
        memory.states.append((inp_graph_1, inp_graph_2))
        memory.actions.append(actions)
        memory.logprobs.append(action_logits)
        return actions

    def evaluate(self, inp_graph_1: List, inp_graph_2: List, action):
        """``ActorCritic.evaluate`` (ged_ppo_bihyb_train.py:122-126) -> (log-probs [2, B], state value [B], entropy [B])."""
        diff_feat, graph_feat_1, graph_feat_2 = self.state_encoder(inp_graph_1, inp_graph_2)
        _, action_logits, entropy = self.actor_net(diff_feat, self.state_encoder.last_mask, action)
        return action_logits, self.value_net(graph_feat_1, graph_feat_2), entropy

    @torch.no_grad()
    def propose(self, graphs_1: List, graphs_2: List, act_n_sel: int) -> Tuple[torch.Tensor, ...]:
        """The candidate fan-out of one beam-search step (ged_ppo_bihyb_eval.py:68-79) for a batch of states: the top
        ``act_n_sel`` first nodes and, for each, the top ``act_n_sel`` second nodes.
        Returns acts1 [B,k], probs1 [B,k], acts2 [B,k,k], probs2 [B,k,k]."""
        k = act_n_sel
        diff_feat, _, _ = self.state_encoder(graphs_1, graphs_2)
        mask = self.state_encoder.last_mask
        acts1, probs1 = self.actor_net.select_nodes(diff_feat, mask, k)
        acts2, probs2 = self.actor_net.select_nodes(diff_feat.repeat_interleave(k, dim=0), mask.repeat_interleave(k, dim=0), k,
                                                    prev_act=acts1.reshape(-1))
        return acts1, probs1, acts2.reshape(-1, k, k), probs2.reshape(-1, k, k)
