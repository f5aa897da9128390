"""TEST INFRASTRUCTURE ONLY -- lets the UNMODIFIED reference policy (``src/ged_ppo_bihyb_model.py``,
``src/pyg_graph_models.py``, ``utils/sinkhorn.py``, ``utils/utils.py``) import in the build container, where
``torch_geometric``, ``torch_scatter`` and ``texttable`` are not installed, to generate the golden vectors of the
"next" row f-1 (batched policy forward for beam search).

Two third-party operators are RESTATED here because the reference leaves them to un-vendored dependencies:
  * ``torch_geometric.nn.GCNConv`` (PyG 1.7 semantics, weight stored (in, out)): symmetric normalisation with one self
    loop per node, ``out[c] += d[r]^-1/2 d[c]^-1/2 (x W)[r]`` over the edges incl. duplicates, plus bias;
  * ``torch_scatter.scatter`` with ``reduce='mean' / 'add'`` along dim 0.
Parity for these two is therefore "unpinned" (no reference implementation available offline); everything else the golden
vectors exercise is the reference's own code.
"""
import importlib
import os
import sys
import types

import torch

from oracle import ref_shim


class _GCNConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        torch.nn.init.xavier_uniform_(self.weight)

    def forward(self, x, edge_index):
        n = x.shape[0]
        row, col = edge_index[0], edge_index[1]
        keep = row != col                                   # add_remaining_self_loops: one self loop of weight 1 per node
        loop = torch.arange(n, device=x.device)
        row = torch.cat([row[keep], loop])
        col = torch.cat([col[keep], loop])
        w = torch.ones(row.shape[0], device=x.device, dtype=x.dtype)
        deg = torch.zeros(n, device=x.device, dtype=x.dtype).index_add_(0, col, w)
        dinv = deg.pow(-0.5)
        dinv[torch.isinf(dinv)] = 0
        norm = dinv[row] * w * dinv[col]
        xw = torch.matmul(x, self.weight)
        out = torch.zeros_like(xw).index_add_(0, col, norm.unsqueeze(-1) * xw[row])
        return out + self.bias


def _scatter(src, index, dim=0, dim_size=None, reduce="add"):
    assert dim == 0
    size = int(index.max().item()) + 1 if dim_size is None else dim_size
    out = torch.zeros((size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros(size, dtype=src.dtype, device=src.device).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        out = out / cnt.clamp(min=1).unsqueeze(-1)
    elif reduce not in ("add", "sum"):
        raise NotImplementedError(reduce)
    return out


def _install():
    ref_shim._install_stubs()
    pyg = sys.modules["torch_geometric"]
    if not hasattr(pyg, "nn"):
        nn_mod = types.ModuleType("torch_geometric.nn")
        nn_mod.GCNConv = _GCNConv
        pyg.nn = nn_mod
        sys.modules["torch_geometric.nn"] = nn_mod
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")
        ts.scatter = _scatter
        sys.modules["torch_scatter"] = ts
    if "texttable" not in sys.modules:
        tt = types.ModuleType("texttable")
        tt.Texttable = object
        sys.modules["texttable"] = tt
    try:
        import scipy.spatial.qhull  # noqa: F401  (utils/utils.py imports QhullError from the removed private module)
    except Exception:
        import scipy.spatial
        q = types.ModuleType("scipy.spatial.qhull")
        q.QhullError = getattr(scipy.spatial, "QhullError", Exception)
        sys.modules["scipy.spatial.qhull"] = q


_cached = None


def load_reference_policy():
    """Returns (model_module, train_module-like namespace with ActorCritic) of the unmodified reference."""
    global _cached
    if _cached is not None:
        return _cached
    _install()
    root = ref_shim.REFERENCE_ROOT
    saved = {k: sys.modules.get(k) for k in ("utils", "src")}
    for name in ("utils", "src"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(root, name)]
        sys.modules[name] = pkg
    try:
        model = importlib.import_module("src.ged_ppo_bihyb_model")
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v

    class ActorCritic(torch.nn.Module):                     # same composition as ged_ppo_bihyb_train.py:101-108
        def __init__(self, node_feature_dim, node_output_size, batch_norm, one_hot_degree, gnn_layers):
            super().__init__()
            self.state_encoder = model.GraphEncoder(node_feature_dim, node_output_size, batch_norm, one_hot_degree, gnn_layers)
            self.actor_net = model.ActorNet(node_output_size, batch_norm)
            self.value_net = model.CriticNet(node_output_size, batch_norm)

    _cached = (model, ActorCritic)
    return _cached
