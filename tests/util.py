"""Shared helpers for the parity tests."""
import numpy as np

from tinysql_b200.chunk import FLOAT64, INT64, UINT64, Chunk, Column


def gen_col(rng, tp, n, null_frac=0.2, lo=-(1 << 62), hi=(1 << 62)):
    """Random column in the style of expression/bench_test.go:56-80 defaultGener (20 % NULL)."""
    if tp == FLOAT64:
        vals = rng.uniform(-1e6, 1e6, n)
    elif tp == UINT64:
        vals = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
        if hi <= (1 << 62):
            vals = rng.integers(max(lo, 0), hi, n, dtype=np.int64).astype(np.uint64)
    else:
        vals = rng.integers(lo, hi, n, dtype=np.int64)
    nn = None
    if null_frac > 0:
        nn = rng.random(n) >= null_frac
    return Column(tp, vals, nn)


def canon(chunk):
    """Order-insensitive canonical form: rows sorted with NULLs as (0, value) / (1,) keys; raw bit patterns."""
    n = chunk.num_rows()
    if n == 0:
        return np.zeros((0, 2 * len(chunk.cols)), dtype=np.uint64)
    parts = []
    for c in chunk.cols:
        nn = c.not_null()
        raw = c.raw().copy()
        raw[~nn] = 0
        parts.append(nn.astype(np.uint64))
        parts.append(raw)
    m = np.stack(parts, axis=1)
    order = np.lexsort(m.T[::-1])
    return m[order]


def assert_same_multiset(a, b):
    ca, cb = canon(a), canon(b)
    assert ca.shape == cb.shape, f"row counts differ: {ca.shape} vs {cb.shape}"
    assert np.array_equal(ca, cb)


def assert_same_ordered(a, b):
    assert a.num_rows() == b.num_rows()
    for x, y in zip(a.cols, b.cols):
        nx, ny = x.not_null(), y.not_null()
        assert np.array_equal(nx, ny)
        assert np.array_equal(x.raw()[nx], y.raw()[ny])


def assert_col_equal(got, want, check_null_slots=False):
    """Bit-exact on NULL masks and on the values of non-NULL rows (NULL slots are don't-care, column.go:150-158)."""
    assert got.length == want.length
    g, w = got.not_null(), want.not_null()
    assert np.array_equal(g, w), f"NULL masks differ at {np.nonzero(g != w)[0][:8]}"
    gr, wr = got.raw(), want.raw()
    if check_null_slots:
        assert np.array_equal(gr, wr)
    else:
        bad = np.nonzero(gr[g] != wr[g])[0]
        assert bad.size == 0, f"values differ at non-NULL rows {bad[:8]}"
