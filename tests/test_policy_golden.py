"""CPU: the dense batched policy (ppo_bihyb_b200.policy, row f-1) against golden vectors of the UNMODIFIED reference
policy (oracle/gen_golden_policy.py; GCNConv / scatter restated in the generator's shim -- parity for those two is
unpinned, see oracle/ref_policy_shim.py).  Floating point: tolerance 1e-5 relative (north_star), stated per assert."""
import os

import numpy as np
import pytest
import torch

from ppo_bihyb_b200.graphs import Data
from ppo_bihyb_b200.policy import ActorCritic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy.npz")
K = 3


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def model(gold):
    ac = ActorCritic(34, 16, False, 0, 3)
    sd = {k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("sd/")}
    missing = ac.load_state_dict(sd, strict=True)           # same parameter names as the reference checkpoints
    assert not missing.missing_keys and not missing.unexpected_keys
    return ac.eval()


def _data(x, ei):
    return Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(ei))


def _close(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    tol = 1e-5 * max(1.0, float(np.abs(want).max()))
    assert got.shape == want.shape, what
    assert np.abs(got - want).max() <= tol, f"{what}: {np.abs(got - want).max()} > {tol}"


def test_beam_like_batch_matches_reference(gold, model):
    n = int(gold["c1/num_states"])
    g1 = [_data(gold["c1/x1"], gold[f"c1/ei1_{i}"]) for i in range(n)]
    g2 = [_data(gold["c1/x2"], gold["c1/ei2"]) for _ in range(n)]
    with torch.no_grad():
        diff, f1, f2 = model.state_encoder(g1, g2)
        value = model.value_net(f1, f2)
    _close(diff, gold["c1/diff"], "diff_feat")
    _close(f1, gold["c1/f1"], "global_feat_1")
    _close(f2, gold["c1/f2"], "global_feat_2")
    _close(value, gold["c1/value"], "state value")
    acts1, probs1, acts2, probs2 = model.propose(g1, g2, K)
    assert np.array_equal(acts1.numpy(), gold["c1/acts1"]) and np.array_equal(acts2.numpy(), gold["c1/acts2"])
    _close(probs1, gold["c1/probs1"], "probs1")
    _close(probs2, gold["c1/probs2"], "probs2")


def test_mixed_size_batch_equals_per_pair_reference(gold, model):
    """All pairs in ONE padded batch (what the batched beam search does) == the reference run pair by pair."""
    P = int(gold["c2/num_pairs"])
    g1 = [_data(gold[f"c2/{p}/x1"], gold[f"c2/{p}/ei1"]) for p in range(P)]
    g2 = [_data(gold[f"c2/{p}/x2"], gold[f"c2/{p}/ei2"]) for p in range(P)]
    with torch.no_grad():
        diff, f1, f2 = model.state_encoder(g1, g2)
    acts1, probs1, acts2, probs2 = model.propose(g1, g2, K)
    for p in range(P):
        n1 = gold[f"c2/{p}/x1"].shape[0]
        _close(diff[p, :n1], gold[f"c2/{p}/diff"][0], f"diff pair {p}")
        assert float(diff[p, n1:].abs().max()) == 0.0 if diff.shape[1] > n1 else True
        _close(f1[p], gold[f"c2/{p}/f1"][0], f"f1 pair {p}")
        _close(f2[p], gold[f"c2/{p}/f2"][0], f"f2 pair {p}")
        # top-k probabilities agree to 1e-5; node ids agree except where two probabilities tie to within that tolerance
        # (a different padded batch shape changes fp32 rounding in the last bit, which can swap exact ties)
        _close(probs1[p], gold[f"c2/{p}/probs1"][0], f"probs1 pair {p}")
        want1, got1 = gold[f"c2/{p}/acts1"][0], acts1[p].numpy()
        assert set(want1.tolist()) == set(got1.tolist()), p
        for pos in np.nonzero(want1 != got1)[0]:
            assert abs(gold[f"c2/{p}/probs1"][0][pos] - float(probs1[p][list(got1).index(want1[pos])])) <= 1e-5
        for j, a in enumerate(want1):                                           # second-node proposals, aligned by first node
            jj = list(got1).index(a)
            _close(probs2[p, jj], gold[f"c2/{p}/probs2"][0][j], f"probs2 pair {p}")
            assert set(acts2[p, jj].tolist()) == set(gold[f"c2/{p}/acts2"][0][j].tolist()), (p, j)
        assert int(acts1[p].max()) < n1 and int(acts2[p].max()) < n1          # padded slots are never proposed


@pytest.mark.reference
def test_pretrained_checkpoints_load_and_match_reference_live():
    """Build container only: the reference's shipped GED checkpoints load unchanged and reproduce the reference policy."""
    from oracle import ref_policy_shim, ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present")
    from ppo_bihyb_b200 import synthetic
    _, RefAC = ref_policy_shim.load_reference_policy()
    path = os.path.join(ref_shim.REFERENCE_ROOT, "pretrained", "PPO_ipfp_datasetAIDS-20-30_beam3_ratio0.2117.pt")
    sd = torch.load(path, map_location="cpu")
    ref, ours = RefAC(34, 64, False, 0, 3).eval(), ActorCritic(34, 64, False, 0, 3).eval()
    ref.load_state_dict(sd)
    ours.load_state_dict(sd)
    for g1, g2 in synthetic.make_pairs("aids-20-30", 3, seed_offset=77):
        r1 = [ref_shim.make_data(torch.from_numpy(g1.features()), torch.from_numpy(g1.edge_index()))]
        r2 = [ref_shim.make_data(torch.from_numpy(g2.features()), torch.from_numpy(g2.edge_index()))]
        with torch.no_grad():
            d_ref, _, _ = ref.state_encoder(r1, r2)
            a_ref, p_ref = ref.actor_net._select_node(d_ref, greedy_sel_num=K)
        acts1, probs1, _, _ = ours.propose([Data.from_host(g1)], [Data.from_host(g2)], K)
        assert np.array_equal(acts1.numpy(), a_ref.numpy())
        _close(probs1, p_ref.numpy(), "probs1 (pretrained)")
