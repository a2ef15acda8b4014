"""GPU parity for multi-column join keys and GROUP BY lists (exact key-tuple fold, tinysql_b200/csrc/dict.cu) against
the CPU oracle, which hashes and compares the full key tuple the way the reference does
(executor/hash_table.go:110-141, util/codec/codec.go:363-382, executor/aggregate.go:359-394)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_py as O
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import FLOAT64, INT64, UINT64, Chunk, Column, tq_array
from tinysql_b200.executor import (AGG_AVG, AGG_COUNT, AGG_FIRSTROW, AGG_MAX, AGG_MIN, AGG_SUM, INNER_JOIN, LEFT_OUTER_JOIN,
                                   RIGHT_OUTER_JOIN, HashAggExec, HashJoinExec, MockDataSource)
from util import assert_same_multiset, gen_col

pytestmark = pytest.mark.gpu

SENT = np.int64(np.uint64(0xA5C3F00DDEADBEEF).astype(np.int64))  # the tables' empty marker, as a key value


def run_join(btypes, bcols, ptypes, pcols, bkeys, pkeys, jt=INNER_JOIN, oir=False, selected=None, chunk=1024, batch=0):
    inner, outer = MockDataSource(btypes, bcols, chunk), MockDataSource(ptypes, pcols, chunk)
    filt = None
    if selected is not None:
        state = {"pos": 0}

        def filt(chk):
            lo = state["pos"]
            state["pos"] += chk.num_rows()
            return selected[lo:state["pos"]]
    e = HashJoinExec(outer, inner, pkeys, bkeys, jt, oir, filt, batch)
    e.Open()
    got = e.drain()
    e.Close()
    want = O.hash_join(jt, oir, btypes, bcols, ptypes, pcols, bkeys, pkeys, selected)
    return got, want


@pytest.mark.parametrize("nb,npr", [(0, 0), (0, 50), (50, 0), (1, 1), (1000, 5000), (5000, 100000)])
@pytest.mark.parametrize("jt,oir", [(INNER_JOIN, False), (INNER_JOIN, True), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)])
def test_join_two_key_columns(lib, nb, npr, jt, oir):
    rng = np.random.default_rng(nb * 13 + npr + jt)
    d1, d2 = max(int(nb ** 0.5), 2), max(int(nb ** 0.5) // 2, 2)
    bcols = [gen_col(rng, INT64, nb, 0.05, 0, d1), gen_col(rng, INT64, nb, 0.1), gen_col(rng, INT64, nb, 0.05, -d2, d2)]
    pcols = [gen_col(rng, INT64, npr, 0.05, -d2, d2 + 2), gen_col(rng, FLOAT64, npr, 0.1), gen_col(rng, INT64, npr, 0.05, 0, d1 + 2)]
    got, want = run_join([INT64] * 3, bcols, [INT64, FLOAT64, INT64], pcols, [0, 2], [2, 0], jt, oir)
    assert len(got.cols) == 6
    assert_same_multiset(got, want)


@pytest.mark.parametrize("jt,oir", [(INNER_JOIN, True), (LEFT_OUTER_JOIN, False)])
def test_join_three_key_columns_mixed_types_batches(lib, jt, oir):
    """three key columns (signed / unsigned / double), an outer-side filter, several probe batches, ragged chunks"""
    rng = np.random.default_rng(99 + jt)
    nb, npr = 4000, 60000
    bu = Column(UINT64, rng.integers(0, 12, nb).astype(np.uint64), rng.random(nb) > 0.03)
    bf = Column(FLOAT64, rng.integers(0, 5, nb) * 0.5, rng.random(nb) > 0.03)
    bcols = [gen_col(rng, INT64, nb, 0.03, -6, 6), bu, bf, Column(INT64, np.arange(nb))]
    pi = Column(INT64, rng.integers(-2, 14, npr), rng.random(npr) > 0.03)       # compared with the UNSIGNED build column
    pf = Column(FLOAT64, rng.integers(0, 6, npr) * 0.5, rng.random(npr) > 0.03)
    pcols = [Column(INT64, np.arange(npr)), pf, gen_col(rng, INT64, npr, 0.03, -7, 7), pi]
    sel = (rng.random(npr) < 0.8).astype(np.uint8)
    got, want = run_join([INT64, UINT64, FLOAT64, INT64], bcols, [INT64, FLOAT64, INT64, INT64], pcols, [0, 1, 2], [2, 3, 1], jt, oir,
                         selected=sel, chunk=1000, batch=8192)
    assert_same_multiset(got, want)


def test_join_multi_key_special_values(lib):
    """the empty-marker value, -1 vs 2^64-1 across signedness, NaN / +-0.0 doubles, NULL in either key column"""
    big = np.uint64((1 << 64) - 1)
    bcols = [Column(UINT64, np.array([1, big, 5, np.uint64(SENT), 1, 7], dtype=np.uint64), [True, True, True, True, True, False]),
             Column(FLOAT64, [0.0, 1.0, np.nan, 2.0, 0.0, 3.0]), Column(INT64, [10, 11, 12, 13, 14, 15])]
    pcols = [Column(INT64, [1, -1, 5, SENT, 1, 7, 1], [True] * 6 + [False]),
             Column(FLOAT64, [0.0, 1.0, np.nan, 2.0, -0.0, 3.0, 0.0])]
    for jt, oir in ((INNER_JOIN, False), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)):
        got, want = run_join([UINT64, FLOAT64, INT64], bcols, [INT64, FLOAT64], pcols, [0, 1], [0, 1], jt, oir)
        assert_same_multiset(got, want)
    # same signedness: the big value and the marker value match themselves
    p2 = [Column(UINT64, np.array([big, np.uint64(SENT), 1], dtype=np.uint64)), Column(FLOAT64, [1.0, 2.0, 0.0])]
    got, want = run_join([UINT64, FLOAT64, INT64], bcols, [UINT64, FLOAT64], p2, [0, 1], [0, 1])
    assert_same_multiset(got, want)
    assert got.num_rows() == 4
    # a DOUBLE key column against an integer one never matches
    got, want = run_join([UINT64, FLOAT64, INT64], bcols, [INT64, FLOAT64], pcols, [0, 1], [0, 0], LEFT_OUTER_JOIN)
    assert_same_multiset(got, want)


def test_join_multi_key_partitioned_build(lib):
    """a build side large enough for the partitioned tables, composite key (a, b) unique per row"""
    rng = np.random.default_rng(17)
    nb, npr = 400000, 1500000
    a = rng.integers(0, 700, nb)
    b = rng.permutation(nb)  # makes (a, b) unique
    bcols = [Column(INT64, a), Column(INT64, b), Column(INT64, a * 1000003 + b)]
    pick = rng.integers(0, nb, npr)
    pa, pb = a[pick].copy(), b[pick].copy()
    miss = rng.random(npr) < 0.3
    pb[miss] += nb  # a value the build side never had
    pcols = [Column(INT64, pb), Column(INT64, pa), Column(INT64, np.arange(npr))]
    inner, outer = MockDataSource([INT64] * 3, bcols, 1 << 19), MockDataSource([INT64] * 3, pcols, 1 << 19)
    e = HashJoinExec(outer, inner, [1, 0], [0, 1], INNER_JOIN, True, None, 1 << 19)
    e.Open()
    got = e.drain()
    e.Close()
    assert got.num_rows() == int((~miss).sum())
    # properties: build payload == a*1000003+b of the probe row's key; every matching probe id exactly once
    assert np.array_equal(got.cols[0].values, got.cols[4].values) and np.array_equal(got.cols[1].values, got.cols[3].values)
    assert np.array_equal(got.cols[2].values, got.cols[0].values * 1000003 + got.cols[1].values)
    assert np.array_equal(np.sort(got.cols[5].values), np.nonzero(~miss)[0])


# ------------------------------------------------------------------ GROUP BY a, b[, c]
def run_agg(types, cols, group_by, funcs, chunk=1024, est=0):
    src = MockDataSource(types, cols, chunk)
    e = HashAggExec(src, group_by, funcs, est)
    e.Open()
    got = e.drain()
    e.Close()
    return got


def test_group_by_two_columns_reference_golden(lib):
    # executor/aggregate_test.go:64-66: rows (a,b) where b>0; select count(a) ... group by a, b -> sorted "1","1","1","3"
    rows = [(1, 1), (3, 3), (3, 2), (2, 1), (1, 1), (1, 1)]
    a, b = Column(INT64, [r[0] for r in rows]), Column(INT64, [r[1] for r in rows])
    got = run_agg([INT64, INT64], [a, b], [0, 1], [(AGG_COUNT, 0), (AGG_FIRSTROW, 0), (AGG_FIRSTROW, 1)])
    assert sorted(got.rows()) == [(1, 2, 1), (1, 3, 2), (1, 3, 3), (3, 1, 1)]


@pytest.mark.parametrize("n,d1,d2", [(0, 1, 1), (1, 1, 1), (5000, 7, 9), (200000, 300, 40), (300000, 100000, 3)])
def test_group_by_two_columns(lib, n, d1, d2):
    rng = np.random.default_rng(n + d1)
    k1 = gen_col(rng, INT64, n, 0.05, -d1, d1)
    k2 = gen_col(rng, UINT64, n, 0.05, 0, d2)
    x = Column(FLOAT64, np.floor(rng.random(n) * 4096) / 16, rng.random(n) > 0.1)   # dyadic: float sums exact in any order
    v = gen_col(rng, INT64, n, 0.1, -1000, 1000)
    types, cols = [INT64, FLOAT64, UINT64, INT64], [k1, x, k2, v]
    funcs = [(AGG_FIRSTROW, 0), (AGG_FIRSTROW, 2), (AGG_COUNT, -1), (AGG_COUNT, 1), (AGG_SUM, 1), (AGG_AVG, 1), (AGG_SUM, 3), (AGG_AVG, 3),
             (AGG_MAX, 3), (AGG_MIN, 1)]
    got = run_agg(types, cols, [0, 2], funcs, est=d1 * d2)
    rc, want = O.hash_agg(types, cols, [0, 2], funcs, 3)
    assert rc == 0
    assert_same_multiset(got, want)


def test_group_by_three_columns_growth_and_special_values(lib):
    """three GROUP BY columns (the pair dictionary is exercised), est_groups far too small, the marker value and NULLs as keys"""
    rng = np.random.default_rng(4)
    n = 250000
    k1 = gen_col(rng, INT64, n, 0.02, 0, 50)
    k2 = gen_col(rng, INT64, n, 0.02, 0, 60)
    k3 = Column(FLOAT64, rng.integers(0, 40, n) * 0.25, rng.random(n) > 0.02)
    k1.values[:100] = SENT
    k2.values[50:150] = SENT
    v = Column(INT64, rng.integers(-100, 100, n))
    types, cols = [INT64, INT64, FLOAT64, INT64], [k1, k2, k3, v]
    funcs = [(AGG_FIRSTROW, 2), (AGG_FIRSTROW, 1), (AGG_FIRSTROW, 0), (AGG_SUM, 3), (AGG_COUNT, -1), (AGG_MAX, 3)]
    got = run_agg(types, cols, [0, 1, 2], funcs, est=1)
    rc, want = O.hash_agg(types, cols, [0, 1, 2], funcs, 2)
    assert rc == 0
    assert_same_multiset(got, want)


def test_group_by_two_columns_partial_then_final(lib):
    """Partial1 handles over disjoint halves -> export_partial (key columns first) -> merge_partial into a Final handle"""
    rng = np.random.default_rng(45)
    n = 100000
    k1, k2 = gen_col(rng, INT64, n, 0.03, 0, 60), gen_col(rng, INT64, n, 0.03, -30, 30)
    x = Column(FLOAT64, np.floor(rng.random(n) * 1024) / 8, rng.random(n) > 0.1)
    v = gen_col(rng, INT64, n, 0.1, -500, 500)
    types = [INT64, INT64, FLOAT64, INT64]
    funcs = [(AGG_FIRSTROW, 0), (AGG_FIRSTROW, 1), (AGG_COUNT, -1), (AGG_SUM, 2), (AGG_AVG, 2), (AGG_AVG, 3), (AGG_MAX, 3), (AGG_MIN, 2)]

    def make():
        it, gb = (C.c_int32 * 4)(*types), (C.c_int32 * 2)(0, 1)
        fa = (L.TQAggFunc * len(funcs))(*[L.TQAggFunc(f, a) for f, a in funcs])
        d = L.TQAggDesc(4, it, 2, gb, len(funcs), fa, 4000)
        h = C.c_void_p()
        L.check(lib.tq_agg_create(C.byref(d), C.byref(h)))
        return h, (it, gb, fa)
    final, keep_f = make()
    width = C.c_int32(0)
    L.check(lib.tq_agg_partial_width(final, C.byref(width)))
    assert width.value == 2 + len(funcs) + 2               # two key columns + one per function, AVG twice
    for lo, hi in ((0, n // 3), (n // 3, n)):
        part, keep_p = make()
        cols = [c.slice(lo, hi) for c in (k1, k2, x, v)]
        L.check(lib.tq_agg_put(part, tq_array(cols), L.TQ_MEM_HOST))
        L.check(lib.tq_agg_eof(part))
        out = (L.TQColumn * width.value)()
        rows = C.c_int64(0)
        L.check(lib.tq_agg_export_partial(part, out, C.byref(rows)))
        assert rows.value > 0
        L.check(lib.tq_agg_merge_partial(final, out, L.TQ_MEM_DEVICE))
        L.check(lib.tq_agg_destroy(part))
    L.check(lib.tq_agg_eof(final))
    out_types = []
    for i in range(len(funcs)):
        t = C.c_int32(0)
        L.check(lib.tq_agg_output_type(final, i, C.byref(t)))
        out_types.append(t.value)
    res = [Column.empty(t, 4096) for t in out_types]
    vals, nns = [[] for _ in funcs], [[] for _ in funcs]
    while True:
        arr = tq_array(res, 4096)
        nr, eof = C.c_int64(0), C.c_int32(0)
        L.check(lib.tq_agg_next(final, 4096, arr, C.byref(nr), C.byref(eof)))
        if nr.value == 0:
            break
        for i, c in enumerate(res):
            vals[i].append(c.values[: nr.value].copy())
            nns[i].append(c.not_null()[: nr.value].copy())
    L.check(lib.tq_agg_destroy(final))
    got = Chunk([Column(t, np.concatenate(a), np.concatenate(b)) for t, a, b in zip(out_types, vals, nns)])
    rc, want = O.hash_agg(types, [k1, k2, x, v], [0, 1], funcs, 2)
    assert rc == 0
    assert_same_multiset(got, want)


# ------------------------------------------------------------------ OtherConditions on the device
@pytest.mark.parametrize("jt,oir", [(INNER_JOIN, False), (INNER_JOIN, True), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)])
@pytest.mark.parametrize("nb,npr", [(0, 50), (300, 2000), (5000, 100000), (300000, 900000)])
def test_join_other_conditions(lib, nb, npr, jt, oir):
    rng = np.random.default_rng(nb + npr + jt)
    ndv = max(nb // 3, 2)
    bcols = [gen_col(rng, INT64, nb, 0.05, 0, ndv), gen_col(rng, INT64, nb, 0.1, -50, 50), gen_col(rng, FLOAT64, nb, 0.1)]
    pcols = [gen_col(rng, FLOAT64, npr, 0.1), gen_col(rng, INT64, npr, 0.05, 0, ndv + 2), gen_col(rng, INT64, npr, 0.1, -50, 50)]
    bt, pt = [INT64, INT64, FLOAT64], [FLOAT64, INT64, INT64]
    # output row = lhs ++ rhs; compare the two integer payloads, one double pair, and a constant
    if oir:   # build side is the left child: out = b0 b1 b2 p0 p1 p2
        conds = [(0, 1, 5), (3, 2, 3), (5, 5, None, INT64, 7)]    # b1 < p2 and b2 >= p0 and p2 != 7
    else:     # out = p0 p1 p2 b0 b1 b2
        conds = [(0, 4, 2), (3, 5, 0), (5, 2, None, INT64, 7)]
    inner, outer = MockDataSource(bt, bcols, 1 << 16), MockDataSource(pt, pcols, 1 << 16)
    e = HashJoinExec(outer, inner, [1], [0], jt, oir, None, 1 << 18, other_conditions=conds)
    e.Open()
    got = e.drain()
    e.Close()
    want = O.hash_join(jt, oir, bt, bcols, pt, pcols, [0], [1], None, conds)
    assert_same_multiset(got, want)


@pytest.mark.parametrize("jt,oir", [(LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)])
@pytest.mark.parametrize("nb,npr", [(300, 2000), (300000, 900000)])
def test_join_default_inner_row(lib, nb, npr, jt, oir):
    """defaultInner (joiner.go:139-143): miss rows of an outer join carry PhysicalHashJoin.DefaultValues on the inner side — e.g.
    COUNT -> 0 after the aggregation push-down (rule_aggregation_push_down.go:211-214) — with and without OtherConditions"""
    rng = np.random.default_rng(nb + jt)
    bcols = [gen_col(rng, INT64, nb, 0.05, 0, nb // 2 + 2), gen_col(rng, INT64, nb, 0.1, -50, 50), gen_col(rng, FLOAT64, nb, 0.1)]
    pcols = [gen_col(rng, INT64, npr, 0.05, 0, nb), gen_col(rng, INT64, npr, 0.1, -50, 50)]
    bt, pt = [INT64, INT64, FLOAT64], [INT64, INT64]
    defaults = [None, 0, 2.5]
    for conds in ((), ([(0, 1, 3)] if oir else [(0, 3, 1)])):   # b1 < p1
        inner, outer = MockDataSource(bt, bcols, 1 << 16), MockDataSource(pt, pcols, 1 << 16)
        e = HashJoinExec(outer, inner, [0], [0], jt, oir, None, 1 << 18, other_conditions=conds, default_inner=defaults)
        e.Open()
        got = e.drain()
        e.Close()
        want = O.hash_join(jt, oir, bt, bcols, pt, pcols, [0], [0], None, conds, default_inner=defaults)
        assert_same_multiset(got, want)
