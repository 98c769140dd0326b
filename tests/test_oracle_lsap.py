"""The C restatement of scipy's rectangular LSAP (oracle/lsap.c) against the
installed scipy.optimize.linear_sum_assignment: identical indices on random,
integer-tied, constant and threshold-clamped matrices (the tie patterns the
tracker produces: cost[cost > max_d] = max_d + 1e-5)."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle.lsap_c import linear_sum_assignment_c


def _matrices(seed, count):
    rng = np.random.default_rng(seed)
    for t in range(count):
        nr, nc = rng.integers(1, 48, 2)
        kind = t % 6
        if kind == 0:
            c = rng.random((nr, nc))
        elif kind == 1:
            c = rng.integers(0, 4, (nr, nc)).astype(float)
        elif kind == 2:
            c = np.full((nr, nc), 0.20001)
        elif kind == 3:
            c = rng.random((nr, nc)); c[c > 0.3] = 0.2 + 1e-5
        elif kind == 4:
            c = rng.random((nr, nc)) * 0.25; c[rng.random((nr, nc)) < 0.85] = 0.7 + 1e-5
        else:
            c = rng.random((nr, nc)); c[:, rng.integers(0, nc)] = 0.0
        yield c


def test_lsap_c_matches_scipy():
    for c in _matrices(1, 1200):
        r1, c1 = linear_sum_assignment(c)
        r2, c2 = linear_sum_assignment_c(c)
        np.testing.assert_array_equal(r1, r2)
        np.testing.assert_array_equal(c1, c2)


@pytest.mark.parametrize("shape", [(0, 5), (5, 0), (1, 1), (1, 7), (7, 1)])
def test_lsap_c_edge_shapes(shape):
    c = np.arange(shape[0] * shape[1], dtype=float).reshape(shape)
    r1, c1 = linear_sum_assignment(c)
    r2, c2 = linear_sum_assignment_c(c)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(c1, c2)


def test_lane_arg_min_tree_equals_the_sequential_scan():
    """csrc/tracker.cu `lsap_lane_best`: the lane's winner among its (value, tie key) candidates is taken by a compare
    tree instead of the sequential scan the solver was first written with.  Both pick min value, then max key; keys are
    unique per column (they encode the column's scan position) and inactive columns carry (inf, 0) -- under those
    conditions the order of comparison cannot matter.  Checked here on the rule itself, ties and infinities included."""
    rng = np.random.default_rng(11)

    def sequential(vals, keys):
        bv, bkey, bk = np.inf, 0, -1
        for k, (v, key) in enumerate(zip(vals, keys)):
            if key and (v < bv or (v == bv and key > bkey)):
                bv, bkey, bk = v, key, k
        return bk

    def tree(vals, keys):
        def better(a, b):               # does b beat a
            return vals[b] < vals[a] or (vals[b] == vals[a] and keys[b] > keys[a])
        w01 = 1 if better(0, 1) else 0
        w23 = 3 if better(2, 3) else 2
        w = w23 if better(w01, w23) else w01
        return w if keys[w] else -1

    pool = np.array([0.0, 0.05, 0.2 + 1e-5, 0.2 + 1e-5, 0.7, np.inf])
    for _ in range(20000):
        vals = rng.choice(pool, 4)
        keys = rng.permutation(np.arange(1, 9))[:4] * rng.integers(0, 2, 4)      # unique non-zero keys, 0 = inactive
        vals = np.where(keys == 0, np.inf, vals)
        assert sequential(vals, keys) == tree(vals, keys), (vals, keys)
