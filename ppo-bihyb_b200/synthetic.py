"""Synthetic AIDS-shaped graph pairs.

The GEDLIB AIDS ``.gxl`` files the reference trains on are fetched with ``svn checkout``
(reference ``ged_data/gedlib_dataset.py:114-115``) and are not available offline, so every measurement uses
"molecule-like" graphs with the dataset's shape statistics:

  * a random spanning tree grown with max degree 3, plus ``round(0.08 n)`` ring-closing edges under a degree-4 cap
    (E ~ 1.08 n undirected edges, mean degree ~2.15, max 4 -- the pretrained policy's input is 29 atom labels +
    one-hot degree 0..4 = 34 features, reference ``ged_env.py:49-59``);
  * labels i.i.d. over 29 atom classes with the skew (0.62, 0.15, 0.12, 0.05, 0.06, 0, ...);
  * node counts per configuration (reference ``ged_data/gedlib_dataset.py:78-83``, ``README.md:40``).

Host-side numpy only; the graphs are moved to the GPU by :mod:`graphs`.
"""
from dataclasses import dataclass

import numpy as np

NUM_LABELS = 29
MAX_DEGREE = 4
LABEL_P = np.array([0.62, 0.15, 0.12, 0.05, 0.06] + [0.0] * (NUM_LABELS - 5))
SEED_BASE = 20211206

CONFIGS = {
    # name: (cfg_id, description)
    "aids-20-30": 1,
    "aids-30-50": 2,
    "aids-50+": 3,
}


@dataclass
class HostGraph:
    """One labelled undirected graph on the host: ``adj`` uint8 [n,n] symmetric 0/1 zero-diagonal, ``labels`` int32 [n]."""
    adj: np.ndarray
    labels: np.ndarray

    @property
    def n(self) -> int:
        return int(self.adj.shape[0])

    def edge_index(self) -> np.ndarray:
        """Both directions of every undirected edge, int64 [2, 2E], row-major order (like PyG's ``to_undirected``)."""
        r, c = np.nonzero(self.adj)
        return np.stack([r, c]).astype(np.int64)

    def features(self) -> np.ndarray:
        """fp32 [n, 34]: one-hot atom label (29) followed by one-hot degree 0..4 (reference ``ged_env.py:53-59``)."""
        n = self.n
        x = np.zeros((n, NUM_LABELS + MAX_DEGREE + 1), np.float32)
        x[np.arange(n), self.labels] = 1
        deg = np.minimum(self.adj.sum(1).astype(np.int64), MAX_DEGREE)
        x[np.arange(n), NUM_LABELS + deg] = 1
        return x


def molecule_like(n: int, rng: np.random.Generator) -> HostGraph:
    adj = np.zeros((n, n), np.uint8)
    deg = np.zeros(n, np.int64)
    for v in range(1, n):
        cand = np.nonzero(deg[:v] < 3)[0]
        if cand.size == 0:                       # cannot happen for a tree grown under cap 3, kept for safety
            cand = np.arange(v)
        u = int(cand[rng.integers(cand.size)])
        adj[u, v] = adj[v, u] = 1
        deg[u] += 1
        deg[v] += 1
    extra = int(round(0.08 * n))
    tries = 0
    while extra > 0 and tries < 50 * n:
        tries += 1
        u, v = (int(t) for t in rng.integers(n, size=2))
        if u == v or adj[u, v] or deg[u] >= MAX_DEGREE or deg[v] >= MAX_DEGREE:
            continue
        adj[u, v] = adj[v, u] = 1
        deg[u] += 1
        deg[v] += 1
        extra -= 1
    labels = rng.choice(NUM_LABELS, size=n, p=LABEL_P).astype(np.int32)
    return HostGraph(adj, labels)


def sample_size(config: str, rng: np.random.Generator) -> int:
    if config == "aids-20-30":
        return int(rng.integers(20, 31))
    if config == "aids-30-50":
        return int(rng.integers(30, 51))
    if config == "aids-50+":
        return 50 + min(50, int(np.floor(rng.exponential(9.6))))
    if config.startswith("n="):
        return int(config[2:])
    raise ValueError(f"unknown synthetic config {config!r}")


def make_pairs(config: str, num_pairs: int, seed_offset: int = 0):
    """``num_pairs`` (graph1, graph2) tuples for a named configuration, deterministic in (config, seed_offset)."""
    cfg_id = CONFIGS.get(config, 4)
    rng = np.random.default_rng(SEED_BASE + cfg_id + 1000003 * seed_offset)
    out = []
    for _ in range(num_pairs):
        n1, n2 = sample_size(config, rng), sample_size(config, rng)
        out.append((molecule_like(n1, rng), molecule_like(n2, rng)))
    return out


def make_edits(pairs, edits_per_pair: int, seed_offset: int = 0):
    """Uniform random unordered toggles (u, v), u != v, on graph 1 of every base pair.

    Returns ``(base_idx int32 [P*E], edit_u int32, edit_v int32)`` in base-pair-major order: the candidate fan-out of
    one beam-search step (reference ``ged_ppo_bihyb_eval.py:95-100``)."""
    rng = np.random.default_rng(SEED_BASE + 77 + 1000003 * seed_offset)
    base, eu, ev = [], [], []
    for p, (g1, _) in enumerate(pairs):
        for _ in range(edits_per_pair):
            u = int(rng.integers(g1.n))
            v = int(rng.integers(g1.n - 1))
            v += v >= u
            base.append(p)
            eu.append(u)
            ev.append(v)
    return np.asarray(base, np.int32), np.asarray(eu, np.int32), np.asarray(ev, np.int32)
