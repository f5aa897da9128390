"""Row f-3: GEDLIB reader / tuple store / cache files against goldens produced by the UNMODIFIED reference
(``oracle/gen_golden_dataset.py``: ``ged_data/gedlib_dataset.py`` run over the synthetic raw directory of ``tests/_gxl.py``,
and the reference's shipped ``ged_data/cache/*.pt``)."""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from _gxl import synthetic_raw_dir

dataset = importlib.import_module("ppo-bihyb_b200.dataset")
ged_env = importlib.import_module("ppo-bihyb_b200.ged_env")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    with np.load(os.path.join(GOLD, "dataset_gedlib.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    r = str(tmp_path_factory.mktemp("gedlib"))
    synthetic_raw_dir(r)
    return r


def test_read_gxl_matches_reference_reader(gold, root):
    raw = os.path.join(root, "raw", "AIDS", "data")
    for k in range(int(gold["gxl_count"])):
        nodes, edges, directed = dataset.read_gxl(os.path.join(raw, str(gold[f"gxl{k}_name"])))
        assert not directed
        assert [a["symbol"] for a in nodes] == list(gold[f"gxl{k}_symbols"])       # untrimmed, like networkx holds them
        assert len(nodes) == len(gold[f"gxl{k}_nodes"])
        ref = {tuple(sorted(e)) for e in gold[f"gxl{k}_edges"].tolist()}            # nx.Graph: undirected, de-duplicated
        assert {tuple(sorted(e)) for e in edges.tolist()} == ref


def test_store_matches_reference_process(gold, root):
    """Same graphs kept (size window, unknown atoms, per-size cap, string-ordered file names), same labels, same coalesced
    undirected edge_index, same ``i`` -- bit for bit."""
    ds = dataset.GEDDataset(root, "AIDS-10-16", set="test")
    name = "AIDS-10-16"
    assert len(ds) == int(gold[f"{name}_len"]) and repr(ds) == f"{name}({len(ds)})"
    graphs = [ds[i] for i in range(len(ds))]
    assert [g.num_nodes for g in graphs] == gold[f"{name}_n"].tolist()
    assert [g.i for g in graphs] == gold[f"{name}_i"].tolist()
    assert np.array_equal(np.concatenate([g.x.argmax(1).numpy() for g in graphs]), gold[f"{name}_labels"])
    assert all(g.x.dtype == torch.float32 and g.x.shape[1] == int(gold[f"{name}_x_dim"]) for g in graphs)
    assert np.array_equal(np.concatenate([g.x.sum(1).numpy() for g in graphs]), gold[f"{name}_x_rowsum"])
    assert np.array_equal(np.cumsum([0] + [g.edge_index.shape[1] for g in graphs]), gold[f"{name}_edge_ptr"])
    assert np.array_equal(np.concatenate([g.edge_index.numpy() for g in graphs], 1), gold[f"{name}_edge_index"])
    assert list(ds.norm_ged.shape) == gold[f"{name}_norm_ged_shape"].tolist() and bool((ds.norm_ged == -1).all())
    # second construction reads the processed store instead of the raw files
    assert os.path.exists(os.path.join(root, "processed", f"{name}.ged_b200.npz"))
    again = dataset.GEDDataset(root, name)
    assert all(torch.equal(a.edge_index, b.edge_index) and torch.equal(a.x, b.x) for a, b in zip(again, graphs))


def test_file_order_is_by_number_string(tmp_path):
    for n in (1000, 200, 30, 4):
        dataset.write_gxl(str(tmp_path / f"{n}.gxl"), ["C"], [])
    assert dataset.gedlib_graph_files(str(tmp_path)) == ["1000.gxl", "200.gxl", "30.gxl", "4.gxl"]


def test_slices_concat_and_one_hot_degree(root):
    ds = dataset.GEDDataset(root, "AIDS-10-16")
    q = len(ds) // 4
    train, val = ds[q:], ds[:q]
    assert len(train) + len(val) == len(ds) and train[0].i == q and val[-1].i == q - 1
    assert train.norm_ged.shape == ds.norm_ged.shape                       # indexed by global i
    assert len(list(train + val + val)) == len(ds) + q
    g = ds[5]
    deg = torch.bincount(g.edge_index[0], minlength=g.num_nodes)
    t = dataset.OneHotDegree(int(deg.max()) + 2)
    h = t(g)
    assert h.x.shape == (g.num_nodes, 29 + int(deg.max()) + 3)
    assert torch.equal(h.x[:, :29], g.x) and torch.equal(h.x[:, 29:].argmax(1), deg) and bool((h.x[:, 29:].sum(1) == 1).all())
    assert g.x.shape[1] == 29                                              # the transform does not modify its input
    with pytest.raises(IndexError):
        ds[len(ds)]


def test_process_dataset_attributes(root):
    env = ged_env.GEDenv("ipfp", "AIDS-10-16", dataset_root=root)
    n = len(env.training_graphs) + len(env.val_graphs)
    assert len(env.val_graphs) == n // 4 and len(env.testing_graphs) == n // 4
    assert env.real_data_size == n and env.ori_feature_dim == 29
    assert env.feature_dim == env.training_graphs[0].x.shape[1] > 29
    assert env.training_graphs[0].x.shape[1] == env.testing_graphs[0].x.shape[1]
    with pytest.raises(ValueError):
        ged_env.GEDenv("ipfp", "nope", dataset_root=root)
    with pytest.raises(NotImplementedError):
        ged_env.GEDenv("ipfp", "AIDS700nef", dataset_root=root)


def test_pair_sampling_reproduces_reference_cache_keys():
    """The keys of the reference's shipped cache files ARE its pair sampling (ged_env.py:62-64): seed = rand_id, two distinct
    graph indices per sample.  140 AIDS-20-30 graphs -> 35 test / 105 train graphs; 15 AIDS-50-100 test graphs."""
    cache = json.load(open(os.path.join(GOLD, "cache_files.json")))
    for fname, n_graphs in (("ipfp_10_AIDS-20-30_1.pt", 35), ("ipfp_50_AIDS-20-30_0.pt", 105), ("beam_10_AIDS-50-100_1.pt", 15)):
        _, num, _, rid = fname[:-3].split("_")
        keys = [f"{a},{b}" for a, b in dataset.sample_pairs(n_graphs, int(num), rid)]
        assert list(dict.fromkeys(keys)) == sorted(cache[fname], key=keys.index)


def test_cache_file_round_trip(tmp_path):
    """Files written here load with a plain ``torch.load`` into the reference's ``{ 'i,j': (tensor([ged]), seconds) }``, and the
    reference's shipped values survive load -> save -> load."""
    cache = json.load(open(os.path.join(GOLD, "cache_files.json")))
    ref = cache["ipfp_10_AIDS-20-30_1.pt"]
    path = dataset.cache_path(str(tmp_path), "ipfp", 10, "AIDS-20-30", 1)
    assert os.path.basename(path) == "ipfp_10_AIDS-20-30_1.pt"
    dataset.save_cache(path, {k: (torch.tensor([v[0]]), v[1]) for k, v in ref.items()})
    back = torch.load(path, weights_only=True)                             # tensors, floats, str keys only
    assert list(back) == list(ref)
    for k, (sol, sec) in back.items():
        assert sol.dtype == torch.float32 and list(sol.shape) == ref[k][2] == [1] and str(sol.dtype) == ref[k][3]
        assert sol.item() == ref[k][0] and sec == ref[k][1]
    assert {k: (s.item(), t) for k, (s, t) in dataset.load_cache(path).items()} == {k: (v[0], v[1]) for k, v in ref.items()}


def test_generate_tuples_needs_gpu_or_cache(root, tmp_path):
    """With cache files for every solver nothing is solved, so this runs on CPU: tuples carry the cached values in the reference's
    layout.  (The GPU test in test_gpu_ged_env.py covers the compute-and-write branch.)"""
    env = ged_env.GEDenv("ipfp", "AIDS-10-16", dataset_root=root)
    pairs = dataset.sample_pairs(len(env.testing_graphs), 5, 3)
    for solver, base in (("hungarian", 20.0), ("ipfp", 10.0), ("rrwm", 19.0), ("beam", 9.0)):
        dataset.save_cache(dataset.cache_path(str(tmp_path), solver, 5, "AIDS-10-16", 3),
                           {f"{a},{b}": (torch.tensor([base + k]), 0.5 + k) for k, (a, b) in enumerate(pairs)})
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):                                   # construct_k needs the device: no CPU fallback
            env.generate_tuples(env.testing_graphs, 5, 3, "cpu", cache_dir=str(tmp_path), verbose=False)
        return
    tuples = env.generate_tuples(env.testing_graphs, 5, 3, "cuda", cache_dir=str(tmp_path), verbose=False)
    assert [t[3] for t in tuples] == [10.0 + k for k in range(5)]
    assert all(set(t[4]) == {"hungarian", "ipfp", "rrwm", "beam"} and t[5]["beam"] == 0.5 + k for k, t in enumerate(tuples))
