"""Graph containers and the factor-form pair set the kernels consume.

``Data`` mirrors the three things the reference's hot path touches on a ``torch_geometric.data.Data``
(``x``, ``edge_index``, ``clone()`` / ``num_nodes``; reference ``utils/ged_env.py:110-121,141-148``).  A real PyG
``Data`` is accepted everywhere a ``Data`` is (duck typing).

``PairSet`` is the HBM layout of a set of base pairs: ragged uint8 adjacencies and int32 labels in two pools with
int64 offsets, exactly the ``ged_pairs`` struct of ``include/ged_b200.h``.  It replaces the dense Kronecker matrix K of
``construct_k`` (``utils/ged_env.py:160-228``): K is (n1+1)^2 (n2+1)^2 floats (55 MB at n=60), the factors are
n1^2 + n2^2 bytes + labels (7.4 KB).
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .synthetic import HostGraph


class Data:
    """Minimal stand-in for ``torch_geometric.data.Data`` (only what the GED path reads)."""

    def __init__(self, x: torch.Tensor = None, edge_index: torch.Tensor = None, **kw):
        self.x = x
        self.edge_index = edge_index
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    def clone(self) -> "Data":
        out = Data()
        for k, v in self.__dict__.items():
            setattr(out, k, v.clone() if torch.is_tensor(v) else v)
        return out

    def to(self, device) -> "Data":
        out = Data()
        for k, v in self.__dict__.items():
            setattr(out, k, v.to(device) if torch.is_tensor(v) else v)
        return out

    def __getitem__(self, key):
        return getattr(self, key)

    @staticmethod
    def from_host(g: HostGraph, device="cpu") -> "Data":
        return Data(x=torch.from_numpy(g.features()).to(device), edge_index=torch.from_numpy(g.edge_index()).to(device))


def data_to_host(g, ori_feature_dim: int) -> Tuple[np.ndarray, Optional[np.ndarray], np.ndarray]:
    """``Data`` -> (adjacency uint8 [n,n] with multiplicities, labels int32 [n] or None, features fp32 [n,F]).

    Labels are recovered from the one-hot block ``x[:, :ori_feature_dim]`` that ``node_metric`` compares
    (reference ``ged_env.py:232-233``); if that block is not one-hot the caller must fall back to an explicit
    node-cost matrix (returned labels are None)."""
    key = host_cache_key(g, ori_feature_dim)
    cached = getattr(g, "_ged_host", None) if isinstance(g, Data) else None
    if cached is not None and cached[0] == key:
        return cached[1], cached[2], cached[3]
    x = g.x.detach().to("cpu", torch.float32).numpy()
    ei = g.edge_index.detach().to("cpu").numpy()
    n = x.shape[0]
    adj = np.zeros((n, n), np.int64)
    if ei.size:
        np.add.at(adj, (ei[0], ei[1]), 1)          # duplicate edges are summed, like PyG's to_dense_adj
    feat = x[:, :ori_feature_dim]
    labels = None
    if feat.shape[1] > 0 and np.all((feat == 0) | (feat == 1)) and np.all(feat.sum(1) == 1):
        labels = feat.argmax(1).astype(np.int32)
    out = (np.minimum(adj, 255).astype(np.uint8), labels, feat)
    if isinstance(g, Data):
        # host view, valid while x / edge_index are the same (unmodified) tensors.  The entry keeps the two tensors alive:
        # the key holds their addresses, and an address can only be handed out again once its tensor is freed.
        g._ged_host = (key,) + out + (ei, (g.edge_index, g.x))
    return out


def host_cache_key(g, ori_feature_dim: int):
    """Identity of the tensors a host view was derived from (a new or in-place modified tensor invalidates it)."""
    return (g.edge_index.data_ptr(), g.edge_index._version, tuple(g.edge_index.shape), g.x.data_ptr(), g.x._version, int(ori_feature_dim))


def node_cost_matrix(feat1: np.ndarray, feat2: np.ndarray) -> np.ndarray:
    """``node_metric`` of the AIDS family (reference ``ged_env.py:230-234,243``) for arbitrary feature rows:
    L1 distance between dummy-padded feature rows mapped through [0, 1, 1]."""
    n1, n2 = feat1.shape[0], feat2.shape[0]
    f1 = np.concatenate([feat1, np.zeros((1, feat1.shape[1]), feat1.dtype)], 0)
    f2 = np.concatenate([feat2, np.zeros((1, feat2.shape[1]), feat2.dtype)], 0)
    enc = np.abs(f1[:, None, :] - f2[None, :, :]).sum(-1).astype(np.int64)
    if enc.max(initial=0) > 2 or enc.min(initial=0) < 0:
        raise IndexError("node_metric: encoding outside the reference's 3-entry mapping table")
    return np.asarray([0.0, 1.0, 1.0], np.float32)[enc].reshape(n1 + 1, n2 + 1)


class PairSet:
    """P base pairs in factor form, resident on one CUDA device (see ``ged_pairs`` in include/ged_b200.h)."""

    def __init__(self, n1, n2, adj1, adj1_off, adj2, adj2_off, lab1, lab1_off, lab2, lab2_off,
                 node_cost, node_cost_off, max_n1: int, max_n2: int, host_n1: np.ndarray, host_n2: np.ndarray):
        self.n1, self.n2 = n1, n2
        self.adj1, self.adj1_off, self.adj2, self.adj2_off = adj1, adj1_off, adj2, adj2_off
        self.lab1, self.lab1_off, self.lab2, self.lab2_off = lab1, lab1_off, lab2, lab2_off
        self.node_cost, self.node_cost_off = node_cost, node_cost_off
        self.max_n1, self.max_n2 = int(max_n1), int(max_n2)
        self.host_n1, self.host_n2 = host_n1, host_n2          # kept on the host: sizes are needed for shapes
        self.num_pairs = int(len(host_n1))
        self.device = adj1.device

    # ---- construction --------------------------------------------------------------------------------------
    @staticmethod
    def pack_host(pairs: Sequence[Tuple[HostGraph, HostGraph]], node_costs: Optional[List[Optional[np.ndarray]]] = None):
        """Pack to pinned host buffers (one buffer per field).  Returns a dict of CPU tensors + size metadata."""
        P = len(pairs)
        n1 = np.fromiter((g1.n for g1, _ in pairs), np.int32, P)
        n2 = np.fromiter((g2.n for _, g2 in pairs), np.int32, P)
        a1_off = np.zeros(P, np.int64)
        a2_off = np.zeros(P, np.int64)
        l1_off = np.zeros(P, np.int64)
        l2_off = np.zeros(P, np.int64)
        if P:
            a1_off[1:] = np.cumsum(n1.astype(np.int64) ** 2)[:-1]
            a2_off[1:] = np.cumsum(n2.astype(np.int64) ** 2)[:-1]
            l1_off[1:] = np.cumsum(n1.astype(np.int64))[:-1]
            l2_off[1:] = np.cumsum(n2.astype(np.int64))[:-1]
        adj1 = np.concatenate([g1.adj.reshape(-1) for g1, _ in pairs]) if P else np.zeros(0, np.uint8)
        adj2 = np.concatenate([g2.adj.reshape(-1) for _, g2 in pairs]) if P else np.zeros(0, np.uint8)
        use_cost = node_costs is not None and any(c is not None for c in node_costs)
        fields = dict(n1=n1, n2=n2, adj1=adj1.astype(np.uint8, copy=False), adj1_off=a1_off,
                      adj2=adj2.astype(np.uint8, copy=False), adj2_off=a2_off)
        if use_cost:
            mats, off, o = [], np.zeros(P, np.int64), 0
            for p, (g1, g2) in enumerate(pairs):
                c = node_costs[p]
                if c is None:
                    c = (g1.labels[:, None] != g2.labels[None, :]).astype(np.float32)
                    c = np.pad(c, ((0, 1), (0, 1)), constant_values=1.0)
                    c[-1, -1] = 0.0
                c = np.ascontiguousarray(c, np.float32)
                assert c.shape == (g1.n + 1, g2.n + 1)
                off[p] = o
                o += c.size
                mats.append(c.reshape(-1))
            fields.update(node_cost=np.concatenate(mats), node_cost_off=off)
        else:
            fields.update(lab1=np.concatenate([g1.labels for g1, _ in pairs]).astype(np.int32) if P else np.zeros(0, np.int32),
                          lab1_off=l1_off,
                          lab2=np.concatenate([g2.labels for _, g2 in pairs]).astype(np.int32) if P else np.zeros(0, np.int32),
                          lab2_off=l2_off)
        tensors = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in fields.items()}
        if torch.cuda.is_available() and sum(t.numel() * t.element_size() for t in tensors.values()) > (1 << 20):
            tensors = {k: (t.pin_memory() if t.numel() else t) for k, t in tensors.items()}    # pinning small batches costs more than it saves
        meta = dict(max_n1=int(n1.max()) if P else 1, max_n2=int(n2.max()) if P else 1, host_n1=n1, host_n2=n2)
        return tensors, meta

    @staticmethod
    def from_packed(tensors, meta, device) -> "PairSet":
        d = {k: t.to(device, non_blocking=True) for k, t in tensors.items()}
        return PairSet(d["n1"], d["n2"], d["adj1"], d["adj1_off"], d["adj2"], d["adj2_off"],
                       d.get("lab1"), d.get("lab1_off"), d.get("lab2"), d.get("lab2_off"),
                       d.get("node_cost"), d.get("node_cost_off"), meta["max_n1"], meta["max_n2"],
                       meta["host_n1"], meta["host_n2"])

    @staticmethod
    def from_host(pairs: Sequence[Tuple[HostGraph, HostGraph]], device, node_costs=None) -> "PairSet":
        tensors, meta = PairSet.pack_host(pairs, node_costs)
        return PairSet.from_packed(tensors, meta, device)

    @staticmethod
    def from_data(pairs, ori_feature_dim: int, device) -> "PairSet":
        """From (Data, Data) tuples, as the reference's call sites hold them."""
        host, costs = [], []
        for g1, g2 in pairs:
            a1, l1, f1 = data_to_host(g1, ori_feature_dim)
            a2, l2, f2 = data_to_host(g2, ori_feature_dim)
            if l1 is None or l2 is None:
                costs.append(node_cost_matrix(f1, f2))
                l1 = np.zeros(a1.shape[0], np.int32)
                l2 = np.zeros(a2.shape[0], np.int32)
            else:
                costs.append(None)
            host.append((HostGraph(a1, l1), HostGraph(a2, l2)))
        return PairSet.from_host(host, device, costs)

    # ---- C view ---------------------------------------------------------------------------------------------
    def c_struct(self) -> _lib.GedPairs:
        p = _lib.ptr
        return _lib.GedPairs(
            self.num_pairs, p(self.n1, _lib.c_i32p), p(self.n2, _lib.c_i32p),
            p(self.adj1, _lib.c_u8p), p(self.adj1_off, _lib.c_i64p), p(self.adj2, _lib.c_u8p), p(self.adj2_off, _lib.c_i64p),
            p(self.lab1, _lib.c_i32p), p(self.lab1_off, _lib.c_i64p), p(self.lab2, _lib.c_i32p), p(self.lab2_off, _lib.c_i64p),
            p(self.node_cost, _lib.c_f32p), p(self.node_cost_off, _lib.c_i64p), self.max_n1, self.max_n2,
            int(getattr(self, "adj1_ld", 0)), int(getattr(self, "adj2_ld", 0)))

    @property
    def mat_stride(self) -> int:
        return (self.max_n1 + 1) * (self.max_n2 + 1)

    def hbm_bytes(self) -> int:
        """Bytes the factor form occupies in HBM (the compulsory input traffic of a solve)."""
        tot = 0
        for t in (self.n1, self.n2, self.adj1, self.adj1_off, self.adj2, self.adj2_off, self.lab1, self.lab1_off,
                  self.lab2, self.lab2_off, self.node_cost, self.node_cost_off):
            if t is not None:
                tot += t.numel() * t.element_size()
        return tot


class FactoredK:
    """Opaque stand-in for the dense ``ori_k`` tensor the reference threads through ``GEDenv.step``.

    ``construct_k`` returns this instead of the (n1+1)^2 (n2+1)^2 matrix.  It behaves like the tensor where the
    reference's callers touch it (``.squeeze(0)`` at ``ged_env.py:83,141``; ``.device`` at ``eval.py:48``) and can
    be expanded with :meth:`dense` (tests, and callers that really want K)."""

    def __init__(self, pairset: PairSet, index: int = 0, host_pair=None):
        self.pairset = pairset
        self.index = index
        self.host_pair = host_pair          # (HostGraph, HostGraph, node_cost or None) for re-packing into batches

    @property
    def device(self):
        return self.pairset.device

    @property
    def n1(self) -> int:
        return int(self.pairset.host_n1[self.index])

    @property
    def n2(self) -> int:
        return int(self.pairset.host_n2[self.index])

    @property
    def shape(self):
        nk = (self.n1 + 1) * (self.n2 + 1)
        return (nk, nk)

    def squeeze(self, dim=None):
        return self

    def to(self, device):
        return self

    def dense(self) -> torch.Tensor:
        """Materialise K exactly as ``construct_k`` would (``utils/ged_env.py:213-226``) -- O(N_K^2) memory."""
        g1, g2, cost = self.host_pair
        return dense_k_from_factors(g1.adj, g2.adj, cost if cost is not None else label_node_cost(g1.labels, g2.labels)).to(self.device)


def label_node_cost(lab1: np.ndarray, lab2: np.ndarray) -> np.ndarray:
    c = (np.asarray(lab1)[:, None] != np.asarray(lab2)[None, :]).astype(np.float32)
    c = np.pad(c, ((0, 1), (0, 1)), constant_values=1.0)
    c[-1, -1] = 0.0
    return c


def dense_k_from_factors(adj1: np.ndarray, adj2: np.ndarray, node_cost: np.ndarray) -> torch.Tensor:
    """K[(i,a),(j,b)] = |A1[i,j]-A2[a,b]| M1[i,j] M2[a,b] off the diagonal, node costs on it (``ged_env.py:213-226``)."""
    n1, n2 = adj1.shape[0], adj2.shape[0]
    A1 = np.zeros((n1 + 1, n1 + 1), np.float32)
    A2 = np.zeros((n2 + 1, n2 + 1), np.float32)
    A1[:n1, :n1] = adj1
    A2[:n2, :n2] = adj2
    M1 = np.ones_like(A1)
    M2 = np.ones_like(A2)
    M1[np.arange(n1), np.arange(n1)] = 0
    M2[np.arange(n2), np.arange(n2)] = 0
    k = np.abs(A1[:, None, :, None] - A2[None, :, None, :]) * (M1[:, None, :, None] * M2[None, :, None, :])
    k = k.reshape((n1 + 1) * (n2 + 1), (n1 + 1) * (n2 + 1))
    np.fill_diagonal(k, np.asarray(node_cost, np.float32).reshape(-1))
    return torch.from_numpy(k)


def factors_from_dense_k(k: torch.Tensor, n1: int, n2: int):
    """Recover (adj1, adj2, node_cost) from a dense K in ``construct_k`` form:
    A1[i,j] = K[(i,n2),(j,n2)], A2[a,b] = K[(n1,a),(n1,b)], node costs = diag(K)."""
    R, C = n1 + 1, n2 + 1
    k4 = k.reshape(R, C, R, C)
    a1 = k4[:n1, n2, :n1, n2].detach().to("cpu")
    a2 = k4[n1, :n2, n1, :n2].detach().to("cpu")
    cost = torch.diagonal(k).reshape(R, C).detach().to("cpu", torch.float32).numpy().copy()
    a1 = a1.numpy().copy()
    a2 = a2.numpy().copy()
    np.fill_diagonal(a1, 0)
    np.fill_diagonal(a2, 0)
    return a1.astype(np.uint8), a2.astype(np.uint8), cost
