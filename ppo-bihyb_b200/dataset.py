"""Row f-3: GEDLIB graph reader, factor-form tuple store and solver-result cache files.

Reference interfaces mirrored (file:line under ``/root/reference``):

  ``read_gxl``                         ``ged_data/gedlib_dataset.py:189-222``  (GXL -> node symbols + edges)
  ``GEDDataset`` (gedlib flavour)      ``ged_data/gedlib_dataset.py:17-178``   (size windows, 25-graphs-per-size cap, 29 atom types)
  ``OneHotDegree``                     ``utils/ged_env.py:54-58`` (torch_geometric.transforms.OneHotDegree, used with cat=True)
  ``GEDenv.process_dataset``           ``utils/ged_env.py:25-59``   -> :func:`process_dataset`
  ``GEDenv.generate_tuples``           ``utils/ged_env.py:61-108``  -> :func:`generate_tuples`
  cache files                          ``ged_data/cache/{solver}_{num_samples}_{dataset}_{rand_id}.pt`` =
                                       ``torch.save({ 'i,j': (tensor([ged]), seconds) })``

What is different on purpose
  * no torch_geometric / networkx / svn: GXL is parsed with ``xml.etree`` straight into ``(labels, undirected edge list)``; the
    processed file is a small ``.npz`` in factor form (CSR-like edge lists + labels), not a pickled PyG ``(data, slices)``.
  * ``generate_tuples`` never builds the dense K (55 MB per pair at n=60, ``ged_env.py:83``): each tuple carries a
    :class:`FactoredK`.  Missing cache files are filled by ONE ``solve_feasible_ged_batch`` launch over all sampled pairs
    (the reference loops pair by pair), and written back in the reference's format, so files are interchangeable both ways.
  * solvers this library does not implement (``'beam'`` = A*-beam in Cython) are taken from cache files when present and
    left out of the tuples otherwise.
"""
import glob
import os
import os.path as osp
import random
import re
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .graphs import Data

# atom vocabulary of the AIDS graphs, in the reference's order (gedlib_dataset.py:87-91) -- the one-hot column index IS the label
AIDS_TYPES = ['O', 'S', 'C', 'N', 'Cl', 'Br', 'B', 'Si', 'Hg', 'I', 'Bi', 'P', 'F', 'Cu', 'Ho', 'Pd', 'Ru', 'Pt', 'Sn', 'Li', 'Ga',
              'Tb', 'As', 'Co', 'Pb', 'Sb', 'Se', 'Ni', 'Te']

# size windows of the GEDLIB subsets (gedlib_dataset.py:54-85)
GEDLIB_SUBSETS = {
    'AIDS-10-16':  dict(data_dir='AIDS/data', min_nodes=10, max_nodes=16, count=4),
    'AIDS-20':     dict(data_dir='AIDS/data', min_nodes=20, max_nodes=20, count=25),
    'AIDS-20-30':  dict(data_dir='AIDS/data', min_nodes=20, max_nodes=30, count=25),
    'AIDS-30-50':  dict(data_dir='AIDS/data', min_nodes=30, max_nodes=50, count=25),
    'AIDS-50-100': dict(data_dir='AIDS/data', min_nodes=50, max_nodes=100, count=25),
}
SIMGNN_DATASETS = ('AIDS700nef', 'LINUX', 'ALKANE', 'IMDBMulti')


def read_gxl(gxl_path: str) -> Tuple[List[Dict], np.ndarray, bool]:
    """One GXL file -> (node attribute dicts in file order, edges int64 [E,2] as node indices, directed flag).

    Node indices are file order, which is what the reference ends up with (networkx insertion order + ``relabel_nodes``,
    gedlib_dataset.py:133-135); edge endpoints resolve to the FIRST node carrying that id (``np.where(...)[0][0]``, :212-213)."""
    root = ET.parse(gxl_path).getroot()
    directed = root.find('graph').get('edgemode', None) == 'directed'
    casts = {'double': float, 'float': float, 'string': str, 'int': int}
    index: Dict[str, int] = {}
    nodes: List[Dict] = []
    for node in root.iter('node'):
        attrs = {}
        for a in node.iter('attr'):
            val = a.find('*')
            attrs[a.get('name')] = casts[val.tag.lower()](val.text)
        nid = node.get('id')
        if nid in index:                    # a repeated id updates the attributes of the same graph node
            nodes[index[nid]].update(attrs)
            continue
        index[nid] = len(nodes)
        nodes.append(attrs)
    edges = [(index[e.get('from')], index[e.get('to')]) for e in root.iter('edge')]
    return nodes, np.asarray(edges, np.int64).reshape(-1, 2), directed


def undirected_edge_index(edges: np.ndarray, num_nodes: int) -> np.ndarray:
    """Both directions, duplicates removed, sorted by (row, col) -- ``torch_geometric.utils.to_undirected`` (:146)."""
    if edges.size == 0:
        return np.zeros((2, 0), np.int64)
    both = np.concatenate([edges, edges[:, ::-1]], 0)
    flat = np.unique(both[:, 0] * num_nodes + both[:, 1])
    return np.stack([flat // num_nodes, flat % num_nodes]).astype(np.int64)


class OneHotDegree:
    """``x <- cat([x, one_hot(out-degree, max_degree+1)])`` (torch_geometric.transforms.OneHotDegree as ged_env.py:54 uses it)."""

    def __init__(self, max_degree: int, cat: bool = True):
        self.max_degree, self.cat = int(max_degree), cat

    def __call__(self, data: Data) -> Data:
        n = data.num_nodes
        deg = torch.bincount(data.edge_index[0], minlength=n) if data.edge_index.numel() else torch.zeros(n, dtype=torch.long)
        hot = torch.nn.functional.one_hot(deg, num_classes=self.max_degree + 1).to(torch.float)
        x = data.x
        if x is not None and self.cat:
            x = x.view(-1, 1) if x.dim() == 1 else x
            hot = torch.cat([x, hot.to(x.dtype)], dim=-1)
        out = Data(**{k: v for k, v in data.__dict__.items() if not k.startswith('_')})
        out.x = hot
        return out


class GraphStore:
    """Factor-form graph collection: labels + undirected edge lists, concatenated, with offsets.  This is what is kept on disk
    (``processed/<name>.ged_b200.npz``) and what :class:`GEDDataset` indexes -- a few bytes per node instead of dense tensors."""

    def __init__(self, labels: np.ndarray, node_ptr: np.ndarray, edge_index: np.ndarray, edge_ptr: np.ndarray, num_types: int):
        self.labels, self.node_ptr = labels.astype(np.int32), node_ptr.astype(np.int64)
        self.edge_index, self.edge_ptr = edge_index.astype(np.int64), edge_ptr.astype(np.int64)
        self.num_types = int(num_types)

    def __len__(self) -> int:
        return len(self.node_ptr) - 1

    @staticmethod
    def from_graphs(graphs: Sequence[Tuple[np.ndarray, np.ndarray]], num_types: int) -> "GraphStore":
        node_ptr = np.cumsum([0] + [len(l) for l, _ in graphs])
        edge_ptr = np.cumsum([0] + [e.shape[1] for _, e in graphs])
        labels = np.concatenate([l for l, _ in graphs]) if graphs else np.zeros(0, np.int32)
        ei = np.concatenate([e for _, e in graphs], 1) if graphs else np.zeros((2, 0), np.int64)
        return GraphStore(labels, node_ptr, ei, edge_ptr, num_types)

    def graph(self, i: int) -> Tuple[np.ndarray, np.ndarray]:
        return (self.labels[self.node_ptr[i]:self.node_ptr[i + 1]], self.edge_index[:, self.edge_ptr[i]:self.edge_ptr[i + 1]])

    def save(self, path: str) -> None:
        os.makedirs(osp.dirname(path) or '.', exist_ok=True)
        with open(path, 'wb') as f:         # explicit handle: np.savez would append ".npz" to the name
            np.savez_compressed(f, labels=self.labels, node_ptr=self.node_ptr, edge_index=self.edge_index, edge_ptr=self.edge_ptr,
                                num_types=np.int64(self.num_types))

    @staticmethod
    def load(path: str) -> "GraphStore":
        with np.load(path) as z:
            return GraphStore(z['labels'], z['node_ptr'], z['edge_index'], z['edge_ptr'], int(z['num_types']))


def gedlib_graph_files(raw_dir: str) -> List[str]:
    """``*.gxl`` under ``raw_dir``, ordered like the reference: by the first digit run of the file name, compared AS A STRING
    (gedlib_dataset.py:125-126) -- '1000.gxl' sorts before '200.gxl'.  Ties keep plain name order."""
    names = sorted(osp.basename(p) for p in glob.glob(osp.join(raw_dir, '*.gxl')))
    return sorted(names, key=lambda x: re.findall(r'[0-9]+', x)[0])


def build_gedlib_store(raw_dir: str, name: str, verbose: bool = False) -> GraphStore:
    """The reference's ``GEDDataset.process`` (gedlib_dataset.py:119-178) without PyG: size window, known atom symbols,
    at most ``count`` graphs per node count, kept in file order; graph ``i`` is its position among the kept graphs."""
    spec = GEDLIB_SUBSETS[name]
    per_size = np.zeros(spec['max_nodes'] - spec['min_nodes'] + 1, np.int64)
    kept = []
    for fname in gedlib_graph_files(raw_dir):
        nodes, edges, _ = read_gxl(osp.join(raw_dir, fname))
        n = len(nodes)
        if n < spec['min_nodes'] or n > spec['max_nodes']:
            continue
        symbols = [str(a.get('symbol', '')).strip() for a in nodes]
        if any(s not in AIDS_TYPES for s in symbols):
            if verbose:
                print(f'skipping {fname}: unknown atom symbol')
            continue
        if per_size[n - spec['min_nodes']] >= spec['count']:
            continue
        per_size[n - spec['min_nodes']] += 1
        kept.append((np.asarray([AIDS_TYPES.index(s) for s in symbols], np.int32), undirected_edge_index(edges, n)))
    return GraphStore.from_graphs(kept, len(AIDS_TYPES))


class GEDDataset:
    """Indexable graph set with the attributes the reference reads from its PyG datasets: ``len``, ``ds[i]`` (a transformed
    :class:`Data` with ``x`` one-hot atom types, ``edge_index``, ``i``), slices, ``ds_a + ds_b``, ``.transform``,
    ``.norm_ged`` (all −1: GEDLIB ships no ground truth, gedlib_dataset.py:101), ``.num_features``.

    ``root`` follows the PyG layout the reference uses (``datasets/GEDLIB``): raw graphs under ``root/raw/AIDS/data/*.gxl``,
    processed store under ``root/processed/<name>.ged_b200.npz`` (built on first use).  ``set`` is accepted and ignored, like
    the reference (:93)."""

    def __init__(self, root: str, name: str, set: str = 'train', transform=None, store: Optional[GraphStore] = None,
                 indices: Optional[np.ndarray] = None):
        if name not in GEDLIB_SUBSETS:
            raise AssertionError(f'unknown GEDLIB subset {name}')
        self.root, self.name, self.transform = root, name, transform
        if store is None:
            path = osp.join(root, 'processed', f'{name}.ged_b200.npz')
            if osp.exists(path):
                store = GraphStore.load(path)
            else:
                raw = osp.join(root, 'raw', GEDLIB_SUBSETS[name]['data_dir'])
                if not osp.isdir(raw):
                    raise FileNotFoundError(f'{raw} not found: fetch gedlib/data/datasets/AIDS/data there (no network here; the '
                                            f'reference uses svn, gedlib_dataset.py:113-117)')
                store = build_gedlib_store(raw, name)
                store.save(path)
        self.store = store
        self.indices = np.arange(len(store)) if indices is None else np.asarray(indices, np.int64)
        # indexed by the graphs' global ``i`` (slices share the full matrix, like PyG's dataset copies do)
        self.norm_ged = torch.full((len(store), len(store)), -1, dtype=torch.float)

    def __len__(self) -> int:
        return len(self.indices)

    def get(self, idx: int) -> Data:
        g = int(self.indices[idx])
        labels, ei = self.store.graph(g)
        x = torch.nn.functional.one_hot(torch.from_numpy(labels.astype(np.int64)), num_classes=self.store.num_types).to(torch.float)
        return Data(x=x, edge_index=torch.from_numpy(np.ascontiguousarray(ei)), i=g)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return GEDDataset(self.root, self.name, transform=self.transform, store=self.store, indices=self.indices[idx])
        if isinstance(idx, (int, np.integer)):
            if idx < 0:
                idx += len(self)
            if not 0 <= idx < len(self):
                raise IndexError(idx)
            data = self.get(int(idx))
            return data if self.transform is None else self.transform(data)
        raise TypeError(f'index {idx!r}')

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __add__(self, other):
        return _Concat([self, other])

    @property
    def num_features(self) -> int:
        return int(self[0].x.shape[1])

    num_node_features = num_features

    def __repr__(self):
        return f'{self.name}({len(self)})'


class _Concat:
    def __init__(self, parts):
        self.parts = list(parts)

    def __add__(self, other):
        return _Concat(self.parts + [other])

    def __len__(self):
        return sum(len(p) for p in self.parts)

    def __iter__(self):
        for p in self.parts:
            yield from p


def process_dataset(env, root: Optional[str] = None) -> None:
    """``GEDenv.process_dataset`` (ged_env.py:25-59) for the GEDLIB subsets: the 1/4 : 3/4 split, the global max degree, the
    ``OneHotDegree`` transform and the ``feature_dim`` / ``ori_feature_dim`` / ``nged_matrix`` attributes."""
    if env.dataset in SIMGNN_DATASETS:
        raise NotImplementedError(f'{env.dataset}: the SimGNN datasets are pickled torch_geometric / networkx objects; this '
                                  f'library reads the GEDLIB (.gxl) subsets {sorted(GEDLIB_SUBSETS)}')
    if env.dataset not in GEDLIB_SUBSETS:
        raise ValueError('Unknown dataset name {}'.format(env.dataset))
    all_graphs = GEDDataset(root or 'datasets/GEDLIB', env.dataset, set='test')
    q = len(all_graphs) // 4
    env.training_graphs, env.val_graphs, env.testing_graphs = all_graphs[q:], all_graphs[:q], all_graphs[:q]
    env.nged_matrix = env.training_graphs.norm_ged
    env.real_data_size = env.nged_matrix.size(0)
    max_degree = 0
    for g in env.training_graphs + env.val_graphs + env.testing_graphs:
        if g.edge_index.size(1) > 0:
            max_degree = max(max_degree, int(torch.bincount(g.edge_index[0]).max().item()))
    hot = OneHotDegree(max_degree, cat=env.training_graphs[0].x is not None)
    env.training_graphs.transform = env.val_graphs.transform = env.testing_graphs.transform = hot
    env.feature_dim = env.training_graphs.num_features
    env.ori_feature_dim = env.feature_dim - max_degree - 1


def cache_path(cache_dir: str, solver: str, num_samples: int, dataset: str, rand_id) -> str:
    return osp.join(cache_dir, f'{solver}_{num_samples}_{dataset}_{rand_id}.pt')


def load_cache(path: str, allow_pickle: bool = False) -> Dict[str, Tuple[torch.Tensor, float]]:
    """``{ 'i,j': (tensor([ged]), seconds) }`` as the reference writes it (ged_env.py:70-77).

    The format holds tensors, floats, tuples and a dict only, so the restricted unpickler (``weights_only=True``) reads
    the reference's files and ours; a cache directory is shared state and must not be able to run code on load.
    ``allow_pickle=True`` is the explicit opt-in for a file the restricted loader rejects."""
    return torch.load(path, map_location='cpu', weights_only=not allow_pickle)


def save_cache(path: str, result: Dict[str, Tuple[torch.Tensor, float]]) -> None:
    os.makedirs(osp.dirname(path) or '.', exist_ok=True)
    torch.save({k: (torch.as_tensor(s, dtype=torch.float32).reshape(1).cpu(), float(t)) for k, (s, t) in result.items()}, path)


def sample_pairs(num_graphs: int, num_samples: int, rand_id) -> List[List[int]]:
    """The reference's pair sampling (ged_env.py:62-64): ``random.seed(int(rand_id))`` then ``random.sample(range(n), 2)``."""
    random.seed(int(rand_id))
    return [random.sample(range(num_graphs), 2) for _ in range(num_samples)]


def generate_tuples(env, graphs_dataset, num_samples: int, rand_id, device, cache_dir: str = 'ged_data/cache', verbose: bool = True):
    """``GEDenv.generate_tuples`` (ged_env.py:61-108): ``[(graph1, graph2, ori_k, reference solution, {solver: ged}, {solver: s})]``
    with ``ori_k`` a :class:`FactoredK`.  Cache files are read when present; for solvers this library runs, a missing file is
    computed in one batched launch and written in the reference's format."""
    indices = sample_pairs(len(graphs_dataset), num_samples, rand_id)
    pairs = [(graphs_dataset[a].to(device), graphs_dataset[b].to(device)) for a, b in indices]
    ged_results = {}
    for key in env.available_solvers:
        path = cache_path(cache_dir, key, num_samples, env.dataset, rand_id)
        if osp.exists(path):
            ged_results[key] = load_cache(path)
        elif key in env.native_solvers:
            sols, _, secs = env.solve_feasible_ged_batch([p[0] for p in pairs], [p[1] for p in pairs], key)
            result = {f'{a},{b}': (s, sec) for (a, b), s, sec in zip(indices, sols, secs)}   # sec = the batch's time / pairs
            save_cache(path, result)
            ged_results[key] = load_cache(path)
    if env.solver_type not in ged_results:
        raise FileNotFoundError(f'no result for solver {env.solver_type!r}: not run by this library and no cache file in {cache_dir}')
    solvers = [k for k in env.available_solvers if k in ged_results]
    tuples, nodes = [], 0
    for n, ((a, b), (g1, g2)) in enumerate(zip(indices, pairs)):
        ori_k = env.construct_k(g1, g2)
        label = float(env.nged_matrix[g1['i'], g2['i']]) * 0.5 * (g1.num_nodes + g2.num_nodes) if hasattr(env, 'nged_matrix') else -1.0
        sols = {k: ged_results[k][f'{a},{b}'][0].item() for k in solvers}
        secs = {k: ged_results[k][f'{a},{b}'][1] for k in solvers}
        if label >= 0:
            sols['label'] = label
        if verbose:
            print(f'id {n} ' + '; '.join(f'{k} ged={sols[k]:.2f} time={secs[k]:.2f}' for k in solvers))
        nodes += g1.num_nodes + g2.num_nodes
        tuples.append((g1, g2, ori_k, sols[env.solver_type], sols, secs))
    if verbose:
        print(f'average number of nodes: {nodes / num_samples / 2}')
        for k in solvers:
            vals = [t[4][k] for t in tuples]
            print(f'{k} ged mean={np.mean(vals):.4f} std={np.std(vals):.4f}')
    return tuples


def write_gxl(path: str, symbols: Sequence[str], edges: Sequence[Tuple[int, int]], graph_id: str = 'molid', valence: int = 1) -> None:
    """Write one graph in the GEDLIB AIDS flavour of GXL (ids ``_1.._n``, a ``symbol`` string and a ``chem`` int per node, a
    ``valence`` int per edge) -- the inverse of :func:`read_gxl`, used for fixtures and for exporting edited graphs."""
    lines = ['<?xml version="1.0"?>', '<!DOCTYPE gxl SYSTEM "http://www.gupro.de/GXL/gxl-1.0.dtd">',
             f'<gxl><graph id="{graph_id}" edgeids="false" edgemode="undirected">']
    for k, sym in enumerate(symbols):
        lines.append(f'<node id="_{k + 1}"><attr name="symbol"><string>{sym}  </string></attr>'
                     f'<attr name="chem"><int>{k % 7}</int></attr><attr name="charge"><int>0</int></attr>'
                     f'<attr name="x"><float>{0.5 * k}</float></attr><attr name="y"><float>{-0.25 * k}</float></attr></node>')
    for a, b in edges:
        lines.append(f'<edge from="_{int(a) + 1}" to="_{int(b) + 1}"><attr name="valence"><int>{valence}</int></attr></edge>')
    lines.append('</graph></gxl>')
    os.makedirs(osp.dirname(path) or '.', exist_ok=True)
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')
