"""GPU parity for FLOAT and var-len KEY columns of HashJoinExec, string / FLOAT GROUP BY items and string / FLOAT aggregate
arguments of HashAggExec — util/codec/codec.go:212-240,276-333 (key encoding), :735-743 (HashGroupKey ETString),
executor/aggfuncs/builder.go:119-172 (maxMin4String, maxMin4Float32, firstRow4String, firstRow4Float32) — against the oracle,
through the C-ABI.  Includes the reference benchmark's DEFAULT key set keyIdx {0, 1} = BIGINT + 5 KiB VARSTRING
(executor/benchmark_test.go:352-360,441-452)."""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, UINT64, Column
from tinysql_b200.executor import (AGG_AVG, AGG_COUNT, AGG_FIRSTROW, AGG_MAX, AGG_MIN, AGG_SUM, INNER_JOIN, LEFT_OUTER_JOIN, RIGHT_OUTER_JOIN,
                                   HashAggExec, HashJoinExec, MockDataSource)
from util import assert_same_multiset, gen_col

pytestmark = pytest.mark.gpu


def run_join(btypes, bcols, ptypes, pcols, bkeys, pkeys, jt=INNER_JOIN, oir=False, chunk=1024, batch=0):
    inner, outer = MockDataSource(btypes, bcols, chunk), MockDataSource(ptypes, pcols, chunk)
    e = HashJoinExec(outer, inner, pkeys, bkeys, jt, oir, None, batch)
    e.Open()
    got = e.drain()
    e.Close()
    return got, O.hash_join(jt, oir, btypes, bcols, ptypes, pcols, bkeys, pkeys)


def words(rng, n, vocab, null_frac=0.1):
    idx = rng.integers(0, len(vocab), n)
    return [vocab[i] if rng.random() >= null_frac else None for i in idx]


VOCAB = [b"", b"a", b"ab", b"abc", b"abd", b"b" * 33, b"\x00", b"\x00\x00", b"xyz" * 100, b"q" * 255, b"q" * 256, b"q" * 257]


def test_join_string_key_explicit(lib):
    # "x" == []byte("x"), "x" != "y" (codec_test.go:767-768); NULL keys never match; duplicates keep build insertion order
    b = [Column(BYTES, [b"x", b"y", None, b"x", b""]), Column(INT64, [10, 20, 30, 40, 50])]
    p = [Column(INT64, [1, 2, 3, 4, 5]), Column(BYTES, [b"x", None, b"z", b"", b"y"])]
    for jt, oir in ((INNER_JOIN, False), (INNER_JOIN, True), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)):
        got, want = run_join([BYTES, INT64], b, [INT64, BYTES], p, [0], [1], jt, oir)
        assert got.rows() == want.rows()
    got, _ = run_join([BYTES, INT64], b, [INT64, BYTES], p, [0], [1])
    assert got.rows() == [(1, b"x", b"x", 10), (1, b"x", b"x", 40), (4, b"", b"", 50), (5, b"y", b"y", 20)]


def test_join_float_key_explicit(lib):
    # float32(1.0) == float64(1.0), != float64(1.1) (codec_test.go:764-765): a FLOAT key is compared as float64(f)
    b = [Column(FLOAT32, np.array([1.0, 0.1, 2.5, -0.0], dtype=np.float32)), Column(INT64, [1, 2, 3, 4])]
    p = [Column(FLOAT64, [1.0, 1.1, 0.1, float(np.float32(0.1)), 2.5, 0.0, -0.0]), Column(INT64, [10, 11, 12, 13, 14, 15, 16])]
    got, want = run_join([FLOAT32, INT64], b, [FLOAT64, INT64], p, [0], [0])
    assert got.rows() == want.rows()
    assert [r[1] for r in got.rows()] == [10, 13, 14, 16]   # 0.1 (double) != float64(0.1f); +0.0 != -0.0 (raw bits)
    # FLOAT with FLOAT, and a FLOAT key never equals an integer key (floatFlag vs varintFlag)
    got, want = run_join([FLOAT32, INT64], b, [FLOAT32, INT64], [Column(FLOAT32, np.array([2.5, 7.0], dtype=np.float32)), Column(INT64, [1, 2])], [0], [0])
    assert got.rows() == want.rows() and got.num_rows() == 1
    got, want = run_join([FLOAT32, INT64], b, [INT64, INT64], [Column(INT64, [1, 2]), Column(INT64, [1, 2])], [0], [0], LEFT_OUTER_JOIN)
    assert got.rows() == want.rows() and got.num_rows() == 2


@pytest.mark.parametrize("jt,oir", [(INNER_JOIN, False), (INNER_JOIN, True), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)])
def test_join_string_key_random(lib, jt, oir):
    rng = np.random.default_rng(100 + jt + 10 * oir)
    nb, npr = 4000, 30000
    vocab = VOCAB + [b"w%04d" % i for i in range(3000)]
    bcols = [Column(BYTES, words(rng, nb, vocab)), gen_col(rng, INT64, nb, 0.1), Column(FLOAT32, rng.random(nb).astype(np.float32))]
    pcols = [gen_col(rng, FLOAT64, npr, 0.1), Column(BYTES, words(rng, npr, vocab + [b"never%d" % i for i in range(500)]))]
    got, want = run_join([BYTES, INT64, FLOAT32], bcols, [FLOAT64, BYTES], pcols, [0], [1], jt, oir, chunk=1000, batch=8192)
    assert_same_multiset(got, want)


def test_join_multi_key_with_string_and_float(lib):
    """(BIGINT, VARCHAR, FLOAT) key tuples: per-column dictionaries + exact fold (dict.cuh) over the 8-byte forms"""
    rng = np.random.default_rng(7)
    nb, npr = 5000, 40000
    vocab = [b"s%d" % i for i in range(40)]
    bcols = [gen_col(rng, INT64, nb, 0.05, 0, 30), Column(BYTES, words(rng, nb, vocab, 0.05)),
             Column(FLOAT32, rng.integers(0, 5, nb).astype(np.float32) * np.float32(0.1), rng.random(nb) > 0.05), Column(INT64, np.arange(nb))]
    pcols = [Column(BYTES, words(rng, npr, vocab + [b"zz"], 0.05)), gen_col(rng, UINT64, npr, 0.05, 0, 32),
             Column(FLOAT64, (rng.integers(0, 6, npr).astype(np.float32) * np.float32(0.1)).astype(np.float64), rng.random(npr) > 0.05)]
    for jt, oir in ((INNER_JOIN, False), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)):
        got, want = run_join([INT64, BYTES, FLOAT32, INT64], bcols, [BYTES, UINT64, FLOAT64], pcols, [0, 1, 2], [1, 0, 2], jt, oir, chunk=1024, batch=16384)
        assert want.num_rows() > npr // 10
        assert_same_multiset(got, want)


def test_join_c1_default_key_set(lib):
    """BASELINE config C1 with the reference benchmark's DEFAULT keyIdx {0, 1}: BIGINT k = row AND the 5 KiB VARSTRING column are
    both join keys (executor/benchmark_test.go:352-360,441-452), 1/5 scale.  Every row matches exactly once."""
    n = 20000
    cell = bytes(range(256)) * 20  # 5 KiB
    k = np.arange(n, dtype=np.int64)
    rng = np.random.default_rng(1)
    build = [Column(INT64, rng.permutation(k)), Column(BYTES, [cell] * n)]
    probe = [Column(INT64, k), Column(BYTES, [cell] * n)]
    inner, outer = MockDataSource([INT64, BYTES], build), MockDataSource([INT64, BYTES], probe)
    e = HashJoinExec(outer, inner, [0, 1], [0, 1], INNER_JOIN, False)
    e.Open()
    rows, seen = 0, np.zeros(n, dtype=bool)
    while True:
        c = e.Next()
        m = c.num_rows()
        if m == 0:
            break
        rows += m
        assert np.array_equal(c.cols[0].values, c.cols[2].values)
        seen[c.cols[0].values] = True
        for v in (c.cols[1], c.cols[3]):
            assert np.array_equal(v.offsets, np.arange(m + 1) * len(cell))
            assert np.array_equal(v.data.reshape(m, len(cell)), np.broadcast_to(np.frombuffer(cell, dtype=np.uint8), (m, len(cell))))
    e.Close()
    assert rows == n and seen.all()
    # one differing byte at the end of the 5 KiB string is a different key
    probe2 = [Column(INT64, k[:100]), Column(BYTES, [cell[:-1] + b"\x00"] * 50 + [cell] * 50)]
    got, want = run_join([INT64, BYTES], build, [INT64, BYTES], probe2, [0, 1], [0, 1])
    assert got.num_rows() == want.num_rows() == 50
    assert_same_multiset(got, want)


def run_agg(types, cols, group_by, funcs, chunk=1024, est=0):
    a = HashAggExec(MockDataSource(types, cols, chunk), group_by, funcs, est)
    a.Open()
    got = a.drain()
    a.Close()
    rc, want = O.hash_agg(types, cols, group_by, funcs)
    assert rc == 0
    return got, want


def by_key(chunk, key_cols, float_cols=()):
    out = {}
    for r in chunk.rows():
        out[tuple(r[c] for c in key_cols)] = r
    return out


def assert_agg_equal(got, want, key_cols, approx_cols=()):
    assert got.num_rows() == want.num_rows()
    assert [c.tp for c in got.cols] == [c.tp for c in want.cols]
    g, w = by_key(got, key_cols), by_key(want, key_cols)
    assert set(g) == set(w)
    for k, wr in w.items():
        gr = g[k]
        for i, (x, y) in enumerate(zip(gr, wr)):
            if i in approx_cols and x is not None and y is not None:
                assert abs(x - y) <= 1e-9 * max(1.0, abs(y)), (k, i, x, y)   # SUM/AVG(float): 1e-9 relative (north_star)
            else:
                assert x == y, (k, i, x, y)


def test_agg_string_group_by_explicit(lib):
    g = Column(BYTES, [b"a", b"b", None, b"a", b"", b"b", None, b"a"])
    s = Column(BYTES, [b"pear", None, b"kiwi", b"apple", b"fig", b"zoo", b"", b"pea"])
    f = Column(FLOAT32, np.array([1.5, 2.5, 0.25, -1.0, 9.0, 2.25, 7.0, 3.0], dtype=np.float32), [True, True, True, True, False, True, True, True])
    funcs = [(AGG_FIRSTROW, 0), (AGG_COUNT, 1), (AGG_MAX, 1), (AGG_MIN, 1), (AGG_MAX, 2), (AGG_MIN, 2), (AGG_SUM, 2), (AGG_AVG, 2), (AGG_COUNT, -1)]
    got, want = run_agg([BYTES, BYTES, FLOAT32], [g, s, f], [0], funcs)
    assert_agg_equal(got, want, [0], approx_cols=(6, 7))
    rows = by_key(got, [0])
    assert rows[(b"a",)] == (b"a", 3, b"pear", b"apple", 3.0, -1.0, 3.5, 3.5 / 3, 3)
    assert rows[(None,)][:4] == (None, 2, b"kiwi", b"")


def test_agg_string_and_float_random(lib):
    rng = np.random.default_rng(21)
    n = 60000
    vocab = VOCAB + [b"g%03d" % i for i in range(300)]
    g = Column(BYTES, words(rng, n, vocab, 0.05))
    s = Column(BYTES, words(rng, n, [b"v%05d" % i for i in range(5000)] + VOCAB, 0.2))
    f = Column(FLOAT32, (rng.random(n) * 100 - 50).astype(np.float32), rng.random(n) > 0.2)
    k2 = gen_col(rng, INT64, n, 0.1, 0, 4)
    # single string GROUP BY item: key passthrough FIRSTROW + string / FLOAT arguments
    funcs = [(AGG_FIRSTROW, 0), (AGG_COUNT, 1), (AGG_MAX, 1), (AGG_MIN, 1), (AGG_MAX, 2), (AGG_MIN, 2), (AGG_SUM, 2), (AGG_AVG, 2)]
    got, want = run_agg([BYTES, BYTES, FLOAT32, INT64], [g, s, f, k2], [0], funcs, chunk=1000)
    assert_agg_equal(got, want, [0], approx_cols=(6, 7))
    # (string, int) GROUP BY items; FLOAT as a GROUP BY item
    funcs = [(AGG_FIRSTROW, 0), (AGG_FIRSTROW, 3), (AGG_COUNT, -1), (AGG_MAX, 1), (AGG_MIN, 2)]
    got, want = run_agg([BYTES, BYTES, FLOAT32, INT64], [g, s, f, k2], [0, 3], funcs, chunk=777)
    assert_agg_equal(got, want, [0, 1])
    fk = Column(FLOAT32, rng.integers(0, 50, n).astype(np.float32) * np.float32(0.1), rng.random(n) > 0.1)
    funcs = [(AGG_FIRSTROW, 0), (AGG_COUNT, -1), (AGG_MIN, 1)]
    got, want = run_agg([FLOAT32, BYTES], [fk, s], [0], funcs)
    assert_agg_equal(got, want, [0])


def test_agg_scalar_string_max_min_and_empty(lib):
    s = Column(BYTES, [b"m", b"zz", None, b"a", b"zz\x00"])
    got, want = run_agg([BYTES], [s], [], [(AGG_MAX, 0), (AGG_MIN, 0), (AGG_COUNT, 0), (AGG_FIRSTROW, 0)])
    assert got.rows() == want.rows() == [(b"zz\x00", b"a", 4, b"m")]
    e = [Column(BYTES, []), Column(FLOAT32, np.zeros(0, dtype=np.float32))]
    got, want = run_agg([BYTES, FLOAT32], e, [], [(AGG_COUNT, 0), (AGG_MAX, 0), (AGG_MIN, 1)])
    assert got.rows() == want.rows() == [(0, None, None)]
