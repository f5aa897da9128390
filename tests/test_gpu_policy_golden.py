"""GPU: the dense batched policy ON THE CUDA DEVICE -- cuBLAS matmuls and the fused inference Sinkhorn (``ged_sinkhorn_batch``,
``policy.log_sinkhorn``) -- against the golden vectors of the UNMODIFIED reference policy (tests/golden/policy.npz,
oracle/gen_golden_policy.py).  Same assertions as the CPU test: outputs within 1e-5 relative, top-k node ids identical."""
import numpy as np
import pytest
import torch

import test_policy_golden as cpu
from ppo_bihyb_b200.policy import ActorCritic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(cpu.GOLD)


@pytest.fixture(scope="module")
def cuda_model(gold, cuda_device):
    torch.backends.cuda.matmul.allow_tf32 = False
    ac = ActorCritic(34, 16, False, 0, 3)
    ac.load_state_dict({k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("sd/")}, strict=True)
    return ac.to(cuda_device).eval()


def test_beam_like_batch_matches_reference_on_cuda(gold, cuda_model, cuda_device):
    n = int(gold["c1/num_states"])
    g1 = [cpu._data(gold["c1/x1"], gold[f"c1/ei1_{i}"]) for i in range(n)]
    g2 = [cpu._data(gold["c1/x2"], gold["c1/ei2"]) for _ in range(n)]
    with torch.no_grad():                                   # no autograd on a CUDA device: the fused Sinkhorn kernel runs
        diff, f1, f2 = cuda_model.state_encoder(g1, g2)
        value = cuda_model.value_net(f1, f2)
    assert diff.is_cuda
    cpu._close(diff.cpu(), gold["c1/diff"], "diff_feat")
    cpu._close(f1.cpu(), gold["c1/f1"], "global_feat_1")
    cpu._close(f2.cpu(), gold["c1/f2"], "global_feat_2")
    cpu._close(value.cpu(), gold["c1/value"], "state value")
    acts1, probs1, acts2, probs2 = (t.cpu() for t in cuda_model.propose(g1, g2, cpu.K))
    assert np.array_equal(acts1.numpy(), gold["c1/acts1"]) and np.array_equal(acts2.numpy(), gold["c1/acts2"])
    cpu._close(probs1, gold["c1/probs1"], "probs1")
    cpu._close(probs2, gold["c1/probs2"], "probs2")
    # the training-time path (autograd on: the torch statement of the same Sinkhorn) agrees with the fused kernel
    with torch.enable_grad():
        diff_t, _, _ = cuda_model.state_encoder(g1, g2)
    cpu._close(diff_t.detach().cpu(), gold["c1/diff"], "diff_feat (autograd path)")


def test_mixed_size_batch_equals_per_pair_reference_on_cuda(gold, cuda_model):
    P = int(gold["c2/num_pairs"])
    g1 = [cpu._data(gold[f"c2/{p}/x1"], gold[f"c2/{p}/ei1"]) for p in range(P)]
    g2 = [cpu._data(gold[f"c2/{p}/x2"], gold[f"c2/{p}/ei2"]) for p in range(P)]
    with torch.no_grad():
        diff, f1, f2 = (t.cpu() for t in cuda_model.state_encoder(g1, g2))
    acts1, probs1, acts2, probs2 = (t.cpu() for t in cuda_model.propose(g1, g2, cpu.K))
    for p in range(P):
        n1 = gold[f"c2/{p}/x1"].shape[0]
        cpu._close(diff[p, :n1], gold[f"c2/{p}/diff"][0], f"diff pair {p}")
        cpu._close(f1[p], gold[f"c2/{p}/f1"][0], f"f1 pair {p}")
        cpu._close(f2[p], gold[f"c2/{p}/f2"][0], f"f2 pair {p}")
        cpu._close(probs1[p], gold[f"c2/{p}/probs1"][0], f"probs1 pair {p}")
        want1, got1 = gold[f"c2/{p}/acts1"][0], acts1[p].numpy()
        assert set(want1.tolist()) == set(got1.tolist()), p
        for pos in np.nonzero(want1 != got1)[0]:            # ids may swap only where two probabilities tie within the tolerance
            assert abs(gold[f"c2/{p}/probs1"][0][pos] - float(probs1[p][list(got1).index(want1[pos])])) <= 1e-5
        for j, a in enumerate(want1):
            jj = list(got1).index(a)
            cpu._close(probs2[p, jj], gold[f"c2/{p}/probs2"][0][j], f"probs2 pair {p}")
            assert set(acts2[p, jj].tolist()) == set(gold[f"c2/{p}/acts2"][0][j].tolist()), (p, j)
        assert int(acts1[p].max()) < n1 and int(acts2[p].max()) < n1
