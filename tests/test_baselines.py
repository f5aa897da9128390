"""Row f-4, CPU part: the batched RRWM / GA glue of ``ppo_bihyb_b200.baselines`` against goldens of the
UNMODIFIED reference ``rrwm_ged`` / ``ga_ged`` (``oracle/gen_golden_baselines.py``), with the three kernel-backed operations
(K v, LSAPE projection, Hungarian start) supplied by the CPU oracle.  The GPU test (test_gpu_baselines.py) runs the same
comparison with the real kernels."""
import types

import numpy as np
import torch

from oracle import ged_oracle as go
from ppo_bihyb_b200 import baselines
from ppo_bihyb_b200.graphs import label_node_cost

from _golden import load, pairs_of

# fp32 exp/log iterates: the reference itself moves by ~1e-6 between BLAS builds; the comparison is on max-normalised iterates
RTOL_EARLY, RTOL_LATE = 2e-5, 2e-3


class OracleOps(baselines._Ragged):
    """``_Ragged`` with K v / projection / start computed by oracle/ged_oracle.c (test infrastructure)."""

    def __init__(self, pairs):
        self.pairs = pairs
        n1 = np.asarray([g1.n for g1, _ in pairs])
        n2 = np.asarray([g2.n for _, g2 in pairs])
        fake = types.SimpleNamespace(device=torch.device("cpu"), host_n1=n1, host_n2=n2, mat_stride=int(((n1 + 1) * (n2 + 1)).max()))
        baselines._lib.require_cuda, keep = (lambda: None), baselines._lib.require_cuda
        try:
            super().__init__(fake)
        finally:
            baselines._lib.require_cuda = keep
        self.cn = [label_node_cost(g1.labels, g2.labels) for g1, g2 in pairs]

    def _each(self, padded):
        for b, (g1, g2) in enumerate(self.pairs):
            yield b, g1, g2, padded[b, :g1.n + 1, :g2.n + 1].numpy()

    def kv(self, padded):
        out = torch.zeros_like(padded)
        for b, g1, g2, x in self._each(padded):
            out[b, :g1.n + 1, :g2.n + 1] = torch.from_numpy(go.kv(g1.adj, g2.adj, self.cn[b], x))
        return out

    def _match(self, fn):
        m1, m2 = self.Rm - 1, self.Cm - 1
        rm = torch.full((self.P, max(m1, 1)), -1, dtype=torch.int32)
        cm = torch.full((self.P, max(m2, 1)), -1, dtype=torch.int32)
        for b, (g1, g2) in enumerate(self.pairs):
            r, c = fn(b, g1, g2)
            rm[b, :g1.n], cm[b, :g2.n] = torch.from_numpy(r), torch.from_numpy(c)
        return rm, cm

    def sinkhorn(self, s, tau, iters):          # the torch statement of the same passes (the GPU path has a fused kernel)
        return baselines._sinkhorn_log(s, self.valid, tau, iters)

    def project(self, padded_cost):
        rm, cm = self._match(lambda b, g1, g2: go.lsape(padded_cost[b, :g1.n + 1, :g2.n + 1].numpy())[:2])
        return self.dense(rm, cm), rm, cm

    def start(self):
        def hun(b, g1, g2):
            r = go.solve(g1.adj, g2.adj, self.cn[b], solver="hungarian")
            return r["rowmatch"], r["colmatch"]
        return self.dense(*self._match(hun))


def _mats(z, key):
    off = z["mat_off"]
    return [z[key][off[b]:off[b + 1]].reshape(int(z["n1"][b]) + 1, int(z["n2"][b]) + 1) for b in range(len(z["n1"]))]


def _close(got, want, rtol):
    return np.abs(got - want).max() <= rtol * np.abs(want).max()


def test_ragged_pad_unpad_round_trip():
    z = load("baselines")
    ops = OracleOps(pairs_of(z))
    rng = np.random.default_rng(0)
    rag = torch.zeros(ops.P, ops.ms)
    for b in range(ops.P):
        rag[b, :ops.R[b] * ops.C[b]] = torch.from_numpy(rng.random(ops.R[b] * ops.C[b]).astype(np.float32))
    pad = ops.pad(rag)
    assert torch.equal(ops.unpad(pad), rag)
    assert all(torch.equal(pad[b, :ops.R[b], :ops.C[b]].reshape(-1), rag[b, :ops.R[b] * ops.C[b]]) for b in range(ops.P))
    assert float(pad[~ops.valid].abs().sum()) == 0.0


def test_rrwm_matches_reference_iterates_and_projection():
    z = load("baselines")
    pairs = pairs_of(z)
    ops = OracleOps(pairs)
    # the Hungarian start on the row-normalised K (ged_env.py:459) equals the start on K itself
    hun = ops.start()
    for b, want in enumerate(_mats(z, "hun_x")):
        assert np.array_equal(hun[b, :want.shape[0], :want.shape[1]].numpy(), want)
    for key, it, rtol in (("rrwm_v1", 1, RTOL_EARLY), ("rrwm_v5", 5, 10 * RTOL_EARLY), ("rrwm_v", 100, RTOL_LATE)):
        tr = {}
        rm, cm, iters = baselines.rrwm_batch(None, max_iter=it, trace=tr, ops=ops)
        for b, want in enumerate(_mats(z, key)):
            got = tr["v"][b, :want.shape[0], :want.shape[1]].numpy()
            assert _close(got, want, rtol), (key, b, np.abs(got - want).max() / np.abs(want).max())
        assert int(iters.max()) <= it
    x = ops.dense(rm, cm)
    same = 0
    for b, ((g1, g2), want) in enumerate(zip(pairs, _mats(z, "rrwm_x"))):
        got = x[b, :want.shape[0], :want.shape[1]].numpy()
        ged = go.score(g1.adj, g2.adj, ops.cn[b], rm[b, :g1.n].numpy(), cm[b, :g2.n].numpy())
        if np.array_equal(got, want):
            same += 1
            assert ged == float(z["rrwm_ged"][b])
    assert same >= len(pairs) - 1            # a projection can flip only where two entries of v are closer than RTOL_LATE


def test_ga_matches_reference():
    z = load("baselines")
    pairs = pairs_of(z)
    ops = OracleOps(pairs)
    tr = {}
    rm, cm = baselines.ga_batch(None, trace=tr, ops=ops)
    x = ops.dense(rm, cm)
    same = close = 0
    for b, ((g1, g2), want_soft, want_x) in enumerate(zip(pairs, _mats(z, "ga_soft"), _mats(z, "ga_x"))):
        got = tr["kv_soft"][b, :want_x.shape[0], :want_x.shape[1]].numpy()
        close += bool(_close(got, want_soft, RTOL_EARLY))
        if np.array_equal(x[b, :want_x.shape[0], :want_x.shape[1]].numpy(), want_x):
            same += 1
            assert go.score(g1.adj, g2.adj, ops.cn[b], rm[b, :g1.n].numpy(), cm[b, :g2.n].numpy()) == float(z["ga_ged"][b])
    # the annealing loop leaves each temperature level when |dv| < 1e-4: a pair sitting on that threshold takes one sweep more
    # or less than the reference did and lands elsewhere on its (near-binary, tau=0.1) soft iterate; the others agree to 1e-7
    assert close >= len(pairs) - 2 and same >= len(pairs) - 1
