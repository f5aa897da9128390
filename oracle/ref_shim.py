"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference solver for golden-vector generation.

This module is used in the build container only (``/root/reference`` does not exist on the GPU box).
It makes ``/root/reference/utils/ged_env.py`` importable without ``torch_geometric`` and the Cython
``a_star`` side-car by registering two tiny stand-in modules, then hands back the reference module
itself.  Nothing under ``ppo-bihyb_b200/`` may import this file; only ``oracle/gen_golden.py`` and
container-only tests do.

What the stand-ins provide (only what ``ged_env.py:7-11,160-201`` touches):
  * ``torch_geometric.data.Data``  -- attribute bag with ``clone``/``num_nodes``/``to``
  * ``torch_geometric.data.Batch.from_data_list`` -- concatenation with node-offsets and ``batch`` vector
  * ``torch_geometric.utils.to_dense_adj`` / ``to_dense_batch`` / ``degree``
  * ``torch_geometric.transforms.OneHotDegree`` (never called on this path)
  * ``a_star.a_star`` (never called: solver_type is 'ipfp'/'hungarian')
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("PPO_BIHYB_REFERENCE", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "ged_env.py")):
    # the GPU box has no /root/reference: oracle/_ref (oracle/build_ref.py; git-ignored, travels with the snapshot)
    REFERENCE_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "ged_env.py"))


class _Data:
    """Minimal stand-in for ``torch_geometric.data.Data``."""

    def __init__(self, x=None, edge_index=None, **kw):
        self.x = x
        self.edge_index = edge_index
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return int(self.x.shape[0])

    def clone(self):
        out = _Data()
        for k, v in self.__dict__.items():
            setattr(out, k, v.clone() if torch.is_tensor(v) else v)
        return out

    def to(self, device):
        out = _Data()
        for k, v in self.__dict__.items():
            setattr(out, k, v.to(device) if torch.is_tensor(v) else v)
        return out

    def __getitem__(self, key):
        return getattr(self, key)


class _Batch(_Data):
    @staticmethod
    def from_data_list(items):
        xs, eis, bvec = [], [], []
        base = 0
        for g, item in enumerate(items):
            n = item.x.shape[0]
            xs.append(item.x)
            eis.append(item.edge_index + base)
            bvec.append(torch.full((n,), g, dtype=torch.long, device=item.x.device))
            base += n
        out = _Batch(x=torch.cat(xs, 0), edge_index=torch.cat(eis, 1))
        out.batch = torch.cat(bvec, 0)
        out.num_graphs = len(items)
        return out


def _to_dense_batch(x, batch=None):
    if batch is None:
        batch = x.new_zeros(x.shape[0], dtype=torch.long)
    nb = int(batch.max().item()) + 1 if batch.numel() else 1
    counts = torch.bincount(batch, minlength=nb)
    nmax = int(counts.max().item())
    starts = torch.cumsum(counts, 0) - counts
    local = torch.arange(x.shape[0], device=x.device) - starts[batch]
    out = x.new_zeros(nb, nmax, x.shape[-1])
    out[batch, local] = x
    mask = torch.zeros(nb, nmax, dtype=torch.bool, device=x.device)
    mask[batch, local] = True
    return out, mask


def _to_dense_adj(edge_index, batch=None, edge_attr=None):
    if batch is None:
        n = int(edge_index.max().item()) + 1 if edge_index.numel() else 0
        batch = edge_index.new_zeros(n)
    nb = int(batch.max().item()) + 1 if batch.numel() else 1
    counts = torch.bincount(batch, minlength=nb)
    nmax = int(counts.max().item())
    starts = torch.cumsum(counts, 0) - counts
    g = batch[edge_index[0]]
    r = edge_index[0] - starts[g]
    c = edge_index[1] - starts[g]
    vals = torch.ones(edge_index.shape[1], device=edge_index.device) if edge_attr is None else edge_attr
    adj = torch.zeros(nb, nmax, nmax, device=edge_index.device, dtype=vals.dtype)
    adj.index_put_((g, r, c), vals, accumulate=True)   # duplicate edges are summed, as PyG does
    return adj


def _degree(index, num_nodes=None, dtype=None):
    n = int(index.max().item()) + 1 if num_nodes is None else num_nodes
    return torch.bincount(index, minlength=n).to(dtype or torch.float)


def _install_stubs():
    if "torch_geometric" not in sys.modules:
        pyg = types.ModuleType("torch_geometric")
        data = types.ModuleType("torch_geometric.data")
        utils = types.ModuleType("torch_geometric.utils")
        transforms = types.ModuleType("torch_geometric.transforms")
        data.Data, data.Batch = _Data, _Batch
        utils.to_dense_batch, utils.to_dense_adj, utils.degree = _to_dense_batch, _to_dense_adj, _degree
        transforms.OneHotDegree = lambda *a, **k: None
        pyg.data, pyg.utils, pyg.transforms = data, utils, transforms
        sys.modules.update({"torch_geometric": pyg, "torch_geometric.data": data,
                            "torch_geometric.utils": utils, "torch_geometric.transforms": transforms})
    if "a_star" not in sys.modules:
        stub = types.ModuleType("a_star")

        def _unavailable(*a, **k):
            raise RuntimeError("a_star side-car is not part of the IPFP path")
        stub.a_star = _unavailable
        sys.modules["a_star"] = stub


_cached = None


def load_reference():
    """Return the reference's own ``utils.ged_env`` module (unmodified source, read from REFERENCE_ROOT)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    # the reference's `utils` is a namespace dir; load it under a private name so it cannot shadow anything here
    spec_dir = os.path.join(REFERENCE_ROOT, "utils")
    pkg = types.ModuleType("utils")
    pkg.__path__ = [spec_dir]
    saved = sys.modules.get("utils")
    sys.modules["utils"] = pkg
    try:
        mod = importlib.import_module("utils.ged_env")
    finally:
        if saved is not None:
            sys.modules["utils"] = saved
    _cached = mod
    return mod


def make_reference_env(solver_type="ipfp", dataset="AIDS-20-30", ori_feature_dim=29):
    """A reference ``GEDenv`` without ``process_dataset`` (needs the GEDLIB download, `ged_env.py:25-59`)."""
    ref = load_reference()
    env = ref.GEDenv.__new__(ref.GEDenv)
    env.solver_type = solver_type
    env.dataset = dataset
    env.ori_feature_dim = ori_feature_dim
    env.available_solvers = ("hungarian", "ipfp", "rrwm", "beam")
    return env


def make_data(x, edge_index):
    _install_stubs()
    return _Data(x=x, edge_index=edge_index)
