"""Deterministic synthetic GEDLIB-style raw directory (shared by oracle/gen_golden_dataset.py and tests/test_dataset.py)."""
import importlib
import os

import numpy as np

dataset = importlib.import_module("ppo-bihyb_b200.dataset")


def synthetic_raw_dir(root: str, seed: int = 7, files: int = 90, lo: int = 8, hi: int = 18) -> str:
    """``root/raw/AIDS/data/<number>.gxl``: sizes around the AIDS-10-16 window (so the window, the 4-per-size cap and the
    unknown-symbol rule all fire), file numbers of mixed digit counts (string-vs-numeric order matters), duplicate and
    reversed edges, a few isolated nodes and one graph without edges."""
    rng = np.random.default_rng(seed)
    raw = os.path.join(root, "raw", "AIDS", "data")
    numbers = rng.choice(np.arange(1, 40000), size=files, replace=False)
    for k, num in enumerate(numbers):
        n = int(rng.integers(lo, hi + 1))
        syms = [dataset.AIDS_TYPES[int(t)] for t in rng.choice([2, 2, 2, 2, 0, 3, 3, 1, 4, 12, 28], size=n)]
        if k % 11 == 5:
            syms[int(rng.integers(n))] = "Xx"                 # unknown atom -> the whole graph is skipped
        edges = [(i, int(rng.integers(i))) for i in range(1, n) if rng.random() < 0.93]      # tree-ish backbone
        edges += [tuple(int(v) for v in rng.choice(n, 2, replace=False)) for _ in range(int(rng.integers(0, 4)))]
        if edges and k % 5 == 0:
            edges.append(edges[0][::-1])                      # reversed duplicate
            edges.append(edges[-2])                           # plain duplicate
        if k == 3:
            edges = []
        dataset.write_gxl(os.path.join(raw, f"{int(num)}.gxl"), syms, edges, graph_id=str(int(num)))
    return raw
