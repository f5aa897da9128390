"""TEST INFRASTRUCTURE -- generates tests/golden/dataset_gedlib.npz and tests/golden/cache_keys.json by running the UNMODIFIED
reference (``/root/reference/ged_data/gedlib_dataset.py``: ``read_gxl`` + ``GEDDataset.process``) over the synthetic raw
directory of ``tests/_gxl.py``, and by reading the reference's shipped ``ged_data/cache/*.pt`` files.

The reference needs torch_geometric (absent) and numpy<1.24 (``np.int``); the stand-ins below provide exactly what
``gedlib_dataset.py`` touches: ``InMemoryDataset`` (raw/processed paths, ``process`` on construction, ``collate`` = identity),
``Data``, ``to_undirected`` (PyG 1.7: both directions, coalesced = sorted by (row, col), unique).  Run in this container only:

    python oracle/gen_golden_dataset.py
"""
import importlib.util
import json
import os
import os.path as osp
import sys
import tempfile
import types

import numpy as np
import torch

HERE = osp.dirname(osp.abspath(__file__))
ROOT = osp.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, osp.join(ROOT, "tests"))
REF = "/root/reference"


class _Data:
    def __init__(self, x=None, edge_index=None, **kw):
        self.x, self.edge_index = x, edge_index
        for k, v in kw.items():
            setattr(self, k, v)


class _InMemoryDataset:
    def __init__(self, root, transform=None, pre_transform=None, pre_filter=None):
        self.root, self.transform, self.pre_transform, self.pre_filter = root, transform, pre_transform, pre_filter
        self.raw_dir, self.processed_dir = osp.join(root, "raw"), osp.join(root, "processed")
        os.makedirs(self.processed_dir, exist_ok=True)
        if not all(osp.exists(p) for p in self.processed_paths):
            self.process()

    @property
    def raw_paths(self):
        return [osp.join(self.raw_dir, f) for f in self.raw_file_names]

    @property
    def processed_paths(self):
        return [osp.join(self.processed_dir, f) for f in self.processed_file_names]

    @staticmethod
    def collate(data_list):
        return data_list, None

    def __len__(self):
        return len(self.data)


def _to_undirected(edge_index, num_nodes=None):
    row, col = edge_index
    row, col = torch.cat([row, col]), torch.cat([col, row])
    flat = torch.unique(row * num_nodes + col)           # sorted
    return torch.stack([flat // num_nodes, flat % num_nodes])


def load_reference_dataset_module():
    pyg, data, utils = (types.ModuleType(n) for n in ("torch_geometric", "torch_geometric.data", "torch_geometric.utils"))
    data.InMemoryDataset, data.Data = _InMemoryDataset, _Data
    data.download_url = data.extract_zip = data.extract_tar = None
    utils.to_undirected = _to_undirected
    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.data", "torch_geometric.utils")}
    sys.modules.update({"torch_geometric": pyg, "torch_geometric.data": data, "torch_geometric.utils": utils})
    if not hasattr(np, "int"):
        np.int = int                                     # the reference predates numpy 1.24 (gedlib_dataset.py:130)
    try:
        spec = importlib.util.spec_from_file_location("ref_gedlib_dataset", osp.join(REF, "ged_data", "gedlib_dataset.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_gedlib_dataset"] = mod
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def main():
    from _gxl import synthetic_raw_dir
    ref = load_reference_dataset_module()
    torch.serialization.add_safe_globals([_Data])        # the reference torch.load()s what its process() saved (:100)
    out = {}
    with tempfile.TemporaryDirectory() as root:
        raw = synthetic_raw_dir(root)
        # read_gxl, graph by graph: node order, symbols, edge set
        names = sorted(os.listdir(raw))
        probe = names[:6]
        for k, name in enumerate(probe):
            G = ref.read_gxl(osp.join(raw, name))
            out[f"gxl{k}_name"] = np.array(name)
            out[f"gxl{k}_nodes"] = np.array(list(G.nodes()))
            out[f"gxl{k}_symbols"] = np.array([G.nodes[v]["symbol"] for v in G.nodes()])
            ids = {v: j for j, v in enumerate(G.nodes())}
            out[f"gxl{k}_edges"] = np.array([[ids[a], ids[b]] for a, b in G.edges()], np.int64).reshape(-1, 2)
        out["gxl_count"] = np.int64(len(probe))
        for name in ("AIDS-10-16",):
            ds = ref.GEDDataset(root, name, set="test")
            graphs = ds.data
            out[f"{name}_len"] = np.int64(len(graphs))
            out[f"{name}_n"] = np.array([int(g.num_nodes) for g in graphs], np.int64)
            out[f"{name}_i"] = np.array([int(g.i) for g in graphs], np.int64)
            out[f"{name}_labels"] = np.concatenate([g.x.argmax(1).numpy() for g in graphs]).astype(np.int32)
            out[f"{name}_x_rowsum"] = np.concatenate([g.x.sum(1).numpy() for g in graphs])
            out[f"{name}_x_dim"] = np.int64(graphs[0].x.shape[1])
            out[f"{name}_edge_ptr"] = np.cumsum([0] + [g.edge_index.shape[1] for g in graphs]).astype(np.int64)
            out[f"{name}_edge_index"] = np.concatenate([g.edge_index.numpy() for g in graphs], 1).astype(np.int64)
            out[f"{name}_norm_ged_shape"] = np.array(ds.norm_ged.shape, np.int64)
            out[f"{name}_norm_ged_all_minus1"] = np.bool_(bool((ds.norm_ged == -1).all()))
    np.savez_compressed(osp.join(ROOT, "tests", "golden", "dataset_gedlib.npz"), **out)

    # the reference's shipped cache files: keys (= its pair sampling), values and timings
    cache = {}
    for fname in sorted(os.listdir(osp.join(REF, "ged_data", "cache"))):
        d = torch.load(osp.join(REF, "ged_data", "cache", fname), weights_only=False)
        cache[fname] = {k: [float(v[0].item()), float(v[1]), list(v[0].shape), str(v[0].dtype)] for k, v in d.items()}
    with open(osp.join(ROOT, "tests", "golden", "cache_files.json"), "w") as f:
        json.dump(cache, f, indent=0, sort_keys=True)
    print("kept", out["AIDS-10-16_len"], "graphs; cache files", len(cache))


if __name__ == "__main__":
    main()
