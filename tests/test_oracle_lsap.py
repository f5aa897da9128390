"""CPU: the oracle's restatement of scipy.optimize.linear_sum_assignment (third party, un-vendored, no version pinned by
the reference; call site utils/ged_env.py:407) against SciPy itself, tie-heavy cases included."""
import numpy as np
import pytest
import scipy.optimize as opt

from oracle import ged_oracle as go


def _cases(rng, num, nmax):
    for t in range(num):
        nr = int(rng.integers(1, nmax))
        nc = int(rng.integers(nr, nmax + 1))
        kind = t % 5
        if kind == 0:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float64)
        elif kind == 1:
            c = rng.integers(0, 7, (nr, nc)) / 3.0
        elif kind == 2:
            c = rng.random((nr, nc))
        elif kind == 3:
            c = np.zeros((nr, nc))
        else:
            c = rng.integers(-3, 4, (nr, nc)).astype(np.float32).astype(np.float64) * 0.1
        yield c


def test_lsap_equals_scipy_rectangular():
    rng = np.random.default_rng(7)
    for c in _cases(rng, 1500, 24):
        row, col = opt.linear_sum_assignment(c)
        c4r, _ = go.lsap(c)
        assert np.array_equal(row, np.arange(c.shape[0])) and np.array_equal(col, c4r)


def test_lsape_equals_scipy_on_the_expanded_matrix():
    """hungarian_lap's block construction (ged_env.py:333-341) through SciPy, vs the oracle's implicit expansion."""
    rng = np.random.default_rng(8)
    for t in range(600):
        n1, n2 = int(rng.integers(0, 16)), int(rng.integers(0, 16))
        if n1 + n2 == 0:
            continue
        C = (rng.integers(0, 4, (n1 + 1, n2 + 1)) / (1.0 if t % 2 else 3.0)).astype(np.float32)
        big = np.full((n1 + n2, n1 + n2), np.inf)
        big[:n1, :n2] = C[:n1, :n2]
        big[np.arange(n1), n2 + np.arange(n1)] = C[:n1, n2]
        big[n1 + np.arange(n2), np.arange(n2)] = C[n1, :n2]
        big[n1:, n2:] = 0.0
        _, col = opt.linear_sum_assignment(big)
        want_rm = np.where(col[:n1] < n2, col[:n1], n2)
        rm, cm, lb, _ = go.lsape(C)
        assert np.array_equal(rm, want_rm), t
        inserted = np.zeros(n2, bool)
        inserted[col[n1:][col[n1:] < n2]] = True
        assert np.array_equal(cm == n1, inserted)


def test_square_lap_embeds_in_lsape_with_same_assignment():
    """A plain square LAP embedded as LSAPE with +inf insertion/deletion costs (the host `hungarian()` drop-in) picks
    the same optimal assignment as SciPy on the plain matrix, ties included."""
    rng = np.random.default_rng(9)
    for t in range(400):
        n = int(rng.integers(1, 20))
        c = rng.integers(0, 3, (n, n)).astype(np.float32) if t % 2 else rng.random((n, n)).astype(np.float32)
        _, col = opt.linear_sum_assignment(c.astype(np.float64))
        C = np.full((n + 1, n + 1), np.inf, np.float32)
        C[:n, :n] = c
        C[n, n] = 0
        rm, cm, _, _ = go.lsape(C)
        assert np.array_equal(rm, col), t


def test_invalid_and_infeasible_inputs():
    with pytest.raises(ValueError):
        go.lsap(np.array([[np.nan, 1.0], [1.0, 0.0]]))
    with pytest.raises(ValueError):
        go.lsap(np.array([[-np.inf, 1.0], [1.0, 0.0]]))
    with pytest.raises(ValueError):
        go.lsap(np.array([[np.inf, np.inf], [1.0, 0.0]]))
    with pytest.raises(ValueError):
        opt.linear_sum_assignment(np.array([[np.inf, np.inf], [1.0, 0.0]]))


def test_lsape_equals_scipy_at_path_sizes():
    """The sizes the path really runs at (round-1 review: SciPy was only consulted up to n = 24): 60 matrices with n1, n2 in
    [50, 100], 6 with n in [150, 200] -- integer ties, k/3 fractions, uniform -- and real IPFP iteration costs of an
    AIDS-50+ solve: the oracle's implicit expansion == SciPy on the expanded (n1+n2)^2 matrix."""
    from ppo_bihyb_b200 import synthetic
    from ppo_bihyb_b200.graphs import label_node_cost
    rng = np.random.default_rng(31)
    mats = []
    for count, lo, hi in ((60, 50, 100), (6, 150, 200)):
        for t in range(count):
            n1, n2 = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
            c = (rng.integers(0, 3, (n1 + 1, n2 + 1)).astype(np.float32) if t % 3 == 0 else
                 (rng.integers(0, 7, (n1 + 1, n2 + 1)) / 3.0).astype(np.float32) if t % 3 == 1 else
                 rng.random((n1 + 1, n2 + 1)).astype(np.float32))
            mats.append(c)
    g1, g2 = synthetic.make_pairs("aids-50+", 1, seed_offset=19)[0]
    o = go.solve(g1.adj, g2.adj, label_node_cost(g1.labels, g2.labels), trace=True)
    mats += [o["cost"][i] for i in range(0, o["iters"], 5)]
    for t, C in enumerate(mats):
        n1, n2 = C.shape[0] - 1, C.shape[1] - 1
        big = np.full((n1 + n2, n1 + n2), np.inf)
        big[:n1, :n2] = C[:n1, :n2]
        big[np.arange(n1), n2 + np.arange(n1)] = C[:n1, n2]
        big[n1 + np.arange(n2), np.arange(n2)] = C[n1, :n2]
        big[n1:, n2:] = 0.0
        row, col = opt.linear_sum_assignment(big)
        r4c = np.empty(n1 + n2, np.int64)
        r4c[col] = row
        rm, cm, _, _ = go.lsape(C)
        assert np.array_equal(rm, np.where(col[:n1] < n2, col[:n1], n2)), t
        assert np.array_equal(cm, np.where(r4c[:n2] < n1, r4c[:n2], n1)), t
