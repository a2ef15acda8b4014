"""GPU parity for SortExec / TopNExec / MergeJoinExec (csrc/sort.cu) through the C-ABI: the reference's own
known answers (tests/sort_cases.py), then the oracle on random tables — ordered equality (both sides keep child order for
rows that compare equal; the reference's sort.Slice leaves that order open) — and size-independent properties at 2e6 rows."""
import numpy as np
import pytest

import oracle_py as O
from sort_cases import MERGE_CASES, SORT_CASES
from test_oracle_sort_merge import random_table
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, UINT64, Column
from tinysql_b200.executor import INNER_JOIN, LEFT_OUTER_JOIN, RIGHT_OUTER_JOIN, MergeJoinExec, MockDataSource, SortExec, TopNExec
from util import assert_same_multiset, assert_same_ordered, gen_col

pytestmark = pytest.mark.gpu


def run_sort(types, cols, by, off=0, cnt=-1, chunk=1024):
    src = MockDataSource(types, cols, chunk)
    e = SortExec(src, by) if cnt < 0 and off == 0 else TopNExec(src, by, off, cnt)
    e.Open()
    got = e.drain()
    e.Close()
    return got


def run_mj(jt, oir, it, ic, ot, oc, ik, ok, sel=None, chunk=1024, default_inner=None, conds=()):
    flt = None
    if sel is not None:
        sel = np.asarray(sel, dtype=np.uint8)
        pos = [0]

        def flt(chk):   # the outer filter's result for the chunk the executor just fetched
            lo = pos[0]
            pos[0] += chk.num_rows()
            return sel[lo: pos[0]]
    e = MergeJoinExec(MockDataSource(ot, oc, chunk), MockDataSource(it, ic, chunk), ok, ik, jt, oir, flt, default_inner=default_inner, other_conditions=conds)
    e.Open()
    got = e.drain()
    e.Close()
    return got


@pytest.mark.parametrize("case", SORT_CASES, ids=[c[0] for c in SORT_CASES])
def test_sort_reference_goldens(lib, case):
    _, types, cols, by, off, cnt, want = case
    assert run_sort(types, cols, by, off, cnt).rows() == want


@pytest.mark.parametrize("case", MERGE_CASES, ids=[c[0] for c in MERGE_CASES])
def test_merge_join_reference_goldens(lib, case):
    _, jt, oir, it, ic, ot, oc, ik, ok, sel, want = case
    assert run_mj(jt, oir, it, ic, ot, oc, ik, ok, sel).rows() == want


@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 257, 4095, 4096, 4097, 70001])
def test_sort_every_layout_vs_oracle(lib, n):
    rng = np.random.default_rng(n)
    types, cols = random_table(rng, n)
    for by in ([(0, False)], [(4, False)], [(4, True), (0, False)], [(2, True), (1, False), (3, True)], [(3, False), (4, True)], [(5, True)], []):
        assert_same_ordered(run_sort(types, cols, by, chunk=1000), O.sort(types, cols, by))
    assert_same_ordered(run_sort(types, cols, [(4, False), (0, True)], 3, 1000), O.sort(types, cols, [(4, False), (0, True)], 3, 1000))


def test_sort_wide_keys_and_long_strings(lib, n=30000):
    rng = np.random.default_rng(3)
    big = Column(INT64, rng.integers(-(1 << 63), (1 << 63) - 1, n, dtype=np.int64), rng.random(n) > 0.05)
    ubig = Column(UINT64, rng.integers(0, (1 << 64) - 1, n, dtype=np.uint64))
    dbl = Column(FLOAT64, np.concatenate([rng.standard_normal(n - 6) * 1e300, [0.0, -0.0, np.inf, -np.inf, 1e-310, -1e-310]]))
    strs = Column(BYTES, [bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)) if rng.random() > 0.05 else None for _ in range(n)])
    rid = Column(INT64, np.arange(n))
    types, cols = [INT64, UINT64, FLOAT64, BYTES, INT64], [big, ubig, dbl, strs, rid]
    for by in ([(0, False)], [(1, True)], [(2, False)], [(3, False)], [(3, True), (0, False)]):
        assert_same_ordered(run_sort(types, cols, by), O.sort(types, cols, by))


def test_sort_large_properties(lib):
    """2e6 rows: the output is a permutation of the input (row ids exactly once), ordered by the key, ties in child order"""
    rng = np.random.default_rng(11)
    n = 2_000_000
    k = rng.integers(-1000, 1000, n)
    f = rng.random(n)
    types, cols = [INT64, FLOAT64, INT64], [Column(INT64, k), Column(FLOAT64, f), Column(INT64, np.arange(n))]
    got = run_sort(types, cols, [(0, False)], chunk=1 << 16)
    gk, gf, gid = (c.values for c in got.cols)
    assert got.num_rows() == n and np.array_equal(np.sort(gid), np.arange(n))
    assert np.array_equal(gk, k[gid]) and np.array_equal(gf, f[gid])
    assert np.all(np.diff(gk) >= 0)
    same = np.diff(gk) == 0
    assert np.all(np.diff(gid)[same] > 0)           # stable
    got = run_sort(types, cols, [(1, True)], 10, 100000, chunk=1 << 16)      # TopN on a float key, descending
    order = np.argsort(-f, kind="stable")[10:100010]
    assert np.array_equal(got.cols[2].values, order)


@pytest.mark.parametrize("jt,oir", [(INNER_JOIN, False), (INNER_JOIN, True), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)])
def test_merge_join_vs_oracle(lib, jt, oir, ni=30000, no=50000):
    rng = np.random.default_rng(60 + jt * 2 + int(oir))
    it, ot = [INT64, BYTES, INT64, FLOAT32], [BYTES, INT64, FLOAT64]
    words = [b"", b"a", b"ab", b"b", b"ba", b"c" * 9, b"c" * 17] + [b"w%04d" % i + b"x" * (i % 23) for i in range(max(ni // 20, 8))]

    def keys(n):
        ki = gen_col(rng, INT64, n, 0.05, 0, 4000)
        ks = Column(BYTES, [words[i] if rng.random() > 0.05 else None for i in rng.integers(0, len(words), n)])
        return ki, ks
    iki, iks = keys(ni)
    oki, oks = keys(no)
    icols = [iki, iks, Column(INT64, np.arange(ni)), Column(FLOAT32, rng.random(ni).astype(np.float32), rng.random(ni) > 0.1)]
    ocols = [oks, oki, Column(FLOAT64, np.arange(no) * 0.5)]
    isorted, osorted = O.sort(it, icols, [(0, False), (1, False)]), O.sort(ot, ocols, [(1, False), (0, False)])
    sel = (rng.random(no) > 0.2).astype(np.uint8)
    for ik, ok in (([0], [1]), ([0, 1], [1, 0]), ([1], [0])):
        ii = isorted if ik != [1] else O.sort(it, icols, [(1, False)])
        oo = osorted if ik != [1] else O.sort(ot, ocols, [(0, False)])
        got = run_mj(jt, oir, it, ii.cols, ot, oo.cols, ik, ok, sel, chunk=1000)
        assert_same_ordered(got, O.merge_join(jt, oir, it, ii.cols, ot, oo.cols, ik, ok, sel))


def test_merge_join_typed_keys_default_inner_and_unsorted_input(lib):
    rng = np.random.default_rng(5)
    n = 5000
    # FLOAT joined with DOUBLE (both evaluate as float64); BIGINT joined with BIGINT UNSIGNED incl. values >= 2^63 and negatives
    fi = Column(FLOAT32, np.sort(rng.integers(0, 300, n)).astype(np.float32) * np.float32(0.5))
    fo = Column(FLOAT64, np.sort(rng.integers(0, 300, n)) * 0.5)
    got = run_mj(LEFT_OUTER_JOIN, False, [FLOAT32], [fi], [FLOAT64], [fo], [0], [0])
    assert_same_ordered(got, O.merge_join(LEFT_OUTER_JOIN, False, [FLOAT32], [fi], [FLOAT64], [fo], [0], [0]))
    si = Column(INT64, np.sort(np.concatenate([rng.integers(-5, 50, n - 2), [-(1 << 63), (1 << 63) - 1]])))
    uo = Column(UINT64, np.sort(np.concatenate([rng.integers(0, 50, n - 3).astype(np.uint64), np.array([(1 << 63), (1 << 64) - 1, (1 << 63) - 1], dtype=np.uint64)])))
    for jt, oir in ((INNER_JOIN, False), (RIGHT_OUTER_JOIN, True)):
        got = run_mj(jt, oir, [INT64], [si], [UINT64], [uo], [0], [0])
        assert_same_ordered(got, O.merge_join(jt, oir, [INT64], [si], [UINT64], [uo], [0], [0]))
    # defaultInner (PhysicalHashJoin / MergeJoin DefaultValues, joiner.go:139-143): COUNT -> 0 on the inner side of a miss row
    inner = [Column(INT64, [1, 3]), Column(INT64, [10, 30])]
    outer = [Column(INT64, [1, 2, 3, 4])]
    got = run_mj(LEFT_OUTER_JOIN, False, [INT64, INT64], inner, [INT64], outer, [0], [0], default_inner=[None, 0])
    assert got.rows() == O.merge_join(LEFT_OUTER_JOIN, False, [INT64, INT64], inner, [INT64], outer, [0], [0], default_inner=[None, 0]).rows() \
        == [(1, 1, 10), (2, None, 0), (3, 3, 30), (4, None, 0)]
    # an inner child that is not sorted by the key is reported, not joined wrongly
    with pytest.raises(L.TQError) as ei:
        run_mj(INNER_JOIN, False, [INT64], [Column(INT64, [3, 1, 2])], [INT64], [Column(INT64, [1, 2, 3])], [0], [0])
    assert ei.value.status == L.TQ_ERR_STATE


@pytest.mark.parametrize("jt,oir", [(INNER_JOIN, False), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True), (INNER_JOIN, True)])
def test_merge_join_other_conditions(lib, jt, oir, ni=20000, no=30000):
    """OtherConditions inside the joiner (baseJoiner.filter): joined rows that fail are dropped, an outer row whose joined rows all
    fail becomes a miss row (merge_join.go:290-305); compared in order with the oracle and, as a multiset, with the hash join"""
    rng = np.random.default_rng(90 + jt + 3 * int(oir))
    it, ot = [INT64, INT64, FLOAT64], [INT64, UINT64, FLOAT64, INT64]
    ic = [Column(INT64, np.sort(rng.integers(0, ni // 4, ni))), gen_col(rng, INT64, ni, 0.1, -50, 50), Column(FLOAT64, rng.integers(0, 100, ni) * 0.5, rng.random(ni) > 0.1)]
    oc = [Column(INT64, np.sort(rng.integers(0, ni // 3, no))), gen_col(rng, UINT64, no, 0.1, 0, 50), Column(FLOAT64, rng.integers(0, 100, no) * 0.5), Column(INT64, np.arange(no))]
    sel = (rng.random(no) > 0.1).astype(np.uint8)
    n_left = len(it) if oir else len(ot)
    icol = lambda c: c if oir else n_left + c            # output column of inner column c
    ocol = lambda c: n_left + c if oir else c            # ... of outer column c
    for conds in ([(0, icol(1), ocol(1))],                                    # inner.b < outer.u  (BIGINT vs BIGINT UNSIGNED)
                  [(3, icol(2), ocol(2)), (5, icol(1), None, INT64, 7)],      # inner.f >= outer.f and inner.b != 7
                  [(4, ocol(2), None, FLOAT64, 12.5)],                        # a condition on the outer side only
                  [(2, icol(1), None, INT64, 1000)]):                         # never true: every match becomes a miss
        got = run_mj(jt, oir, it, ic, ot, oc, [0], [0], sel, chunk=1000, conds=conds)
        want = O.merge_join(jt, oir, it, ic, ot, oc, [0], [0], sel, conds=conds)
        assert_same_ordered(got, want)
        assert_same_multiset(want, O.hash_join(jt, oir, it, ic, ot, oc, [0], [0], sel, conds=conds))


def test_merge_join_equals_hash_join_at_scale(lib):
    """1e6 x 2e6 rows: the merge join's rows are the hash join's rows (same multiset) and come out in outer order"""
    from tinysql_b200.executor import HashJoinExec
    rng = np.random.default_rng(8)
    nb, npr = 1_000_000, 2_000_000
    bk = np.sort(rng.integers(0, 800_000, nb))
    pk = np.sort(rng.integers(0, 900_000, npr))
    b = [Column(INT64, bk), Column(INT64, np.arange(nb))]
    p = [Column(INT64, pk), Column(INT64, np.arange(npr))]
    mj = run_mj(INNER_JOIN, False, [INT64, INT64], b, [INT64, INT64], p, [0], [0], chunk=1 << 16)
    e = HashJoinExec(MockDataSource([INT64, INT64], p, 1 << 16), MockDataSource([INT64, INT64], b, 1 << 16), [0], [0], INNER_JOIN, False)
    e.Open()
    hj = e.drain()
    e.Close()
    assert mj.num_rows() == hj.num_rows()
    pkk, pid, bkk, bid = (c.values for c in mj.cols)
    assert np.array_equal(pkk, bkk) and np.array_equal(pkk, pk[pid]) and np.array_equal(bkk, bk[bid])
    assert np.all(np.diff(pid) >= 0)                       # outer order
    same = np.diff(pid) == 0
    assert np.all(np.diff(bid)[same] > 0)                  # inner order inside a group
    # multiset equality through a checksum of (outer id, inner id) pairs
    h = lambda c: int(((c.cols[1].values.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (c.cols[3].values.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))).sum(dtype=np.uint64))
    assert h(mj) == h(hj)
