"""CPU suite: the oracle's SortExec / TopNExec / MergeJoinExec restatements (executor/sort.go, executor/merge_join.go) against
the reference's own known answers (tests/sort_cases.py) and against independent Python restatements."""
import numpy as np
import pytest

import oracle_py as O
from sort_cases import MERGE_CASES, SORT_CASES
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, UINT64, Column
from util import gen_col


@pytest.mark.parametrize("case", SORT_CASES, ids=[c[0] for c in SORT_CASES])
def test_sort_reference_goldens(case):
    _, types, cols, by, off, cnt, want = case
    assert O.sort(types, cols, by, off, cnt).rows() == want


@pytest.mark.parametrize("case", MERGE_CASES, ids=[c[0] for c in MERGE_CASES])
def test_merge_join_reference_goldens(case):
    _, jt, oir, it, ic, ot, oc, ik, ok, sel, want = case
    assert O.merge_join(jt, oir, it, ic, ot, oc, ik, ok, sel).rows() == want


def py_sort_key(v, desc):
    """cmpNull + the type's order as a Python sort key (NULL first; everything reversed for Desc)"""
    return (0, 0) if v is None else (1, v)


def random_table(rng, n):
    types = [INT64, UINT64, FLOAT64, FLOAT32, BYTES, INT64]
    cols = [gen_col(rng, INT64, n, 0.1, -5, 6), gen_col(rng, UINT64, n, 0.1, 0, 4),
            Column(FLOAT64, rng.integers(-3, 4, n) * 0.5, rng.random(n) > 0.1),
            Column(FLOAT32, (rng.integers(-3, 4, n) * 0.25).astype(np.float32), rng.random(n) > 0.1),
            Column(BYTES, [rng.choice([b"", b"a", b"ab", b"abc", b"b", b"a\x00", b"abcdefgh", b"abcdefghi", b"abcdefgh\x00"]) if rng.random() > 0.1 else None for _ in range(n)]),
            Column(INT64, np.arange(n))]
    return types, cols


@pytest.mark.parametrize("seed", range(6))
def test_sort_matches_python_restatement(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 400))
    types, cols = random_table(rng, n)
    rows = list(zip(*[c.tolist() for c in cols])) if n else []
    for by in ([(0, False)], [(4, False)], [(4, True), (0, False)], [(2, True), (1, False), (3, True)], [(3, False), (4, True)], []):
        got = O.sort(types, cols, by).rows()
        want = sorted(rows, key=lambda r: r[5])   # child order
        for col, desc in reversed(by):
            # stable sort by one ByItem; Desc = reverse comparison with ties still in child order
            keyed = [(py_sort_key(r[col], desc), r) for r in want]
            groups = sorted(set(k for k, _ in keyed), reverse=bool(desc))
            want = [r for g in groups for k, r in keyed if k == g]
        assert got == want
        # TopN = a window of the same order
        assert O.sort(types, cols, by, 3, 5).rows() == want[3:8]


def py_merge_join(jt, oir, irows, orows, ikeys, okeys, sel):
    out = []
    for oi, orow in enumerate(orows):
        key = tuple(orow[k] for k in okeys)
        matches = []
        if (sel is None or sel[oi]) and None not in key:
            matches = [r for r in irows if None not in tuple(r[k] for k in ikeys) and tuple(r[k] for k in ikeys) == key]
        if matches:
            out += [(r + orow) if oir else (orow + r) for r in matches]
        elif jt != 0:
            pad = (None,) * len(irows[0]) if irows else ()
            out.append((pad + orow) if oir else (orow + pad))
    return out


@pytest.mark.parametrize("seed", range(6))
def test_merge_join_matches_python_restatement(seed):
    rng = np.random.default_rng(50 + seed)
    ni, no = int(rng.integers(1, 200)), int(rng.integers(0, 300))
    it, ot = [INT64, BYTES, INT64], [BYTES, INT64, FLOAT64]
    words = [b"", b"a", b"ab", b"b", b"ba", b"c" * 9]

    def side(n, kcol_int, kcol_str):
        ki = gen_col(rng, INT64, n, 0.1, 0, 8)
        ks = Column(BYTES, [words[i] if rng.random() > 0.1 else None for i in rng.integers(0, len(words), n)])
        return ki, ks
    iki, iks = side(ni, 0, 1)
    oki, oks = side(no, 1, 0)
    icols = [iki, iks, Column(INT64, np.arange(ni))]
    ocols = [oks, oki, Column(FLOAT64, np.arange(no) * 0.5)]
    # children sorted ascending by (int key, string key): NULLs first (cmpNull)
    isorted = O.sort(it, icols, [(0, False), (1, False)])
    osorted = O.sort(ot, ocols, [(1, False), (0, False)])
    sel = (rng.random(no) > 0.2).astype(np.uint8)
    for jt, oir in ((0, False), (0, True), (1, False), (2, True)):
        for keys in (([0], [1]), ([0, 1], [1, 0])):
            got = O.merge_join(jt, oir, it, isorted.cols, ot, osorted.cols, keys[0], keys[1], sel).rows()
            want = py_merge_join(jt, oir, isorted.rows(), osorted.rows(), keys[0], keys[1], sel)
            assert got == want


def test_merge_join_equals_hash_join_as_a_multiset():
    """the planner may pick either algorithm for the same equi-join (exhaust_physical_plans.go:281-295)"""
    from util import assert_same_multiset
    rng = np.random.default_rng(7)
    nb, npr = 300, 1000
    b = [gen_col(rng, INT64, nb, 0.1, 0, 50), Column(INT64, np.arange(nb))]
    p = [gen_col(rng, INT64, npr, 0.1, 0, 60), Column(FLOAT64, rng.random(npr))]
    bs, ps = O.sort([INT64, INT64], b, [(0, False)]), O.sort([INT64, FLOAT64], p, [(0, False)])
    for jt, oir in ((0, False), (1, False), (2, True)):
        mj = O.merge_join(jt, oir, [INT64, INT64], bs.cols, [INT64, FLOAT64], ps.cols, [0], [0])
        hj = O.hash_join(jt, oir, [INT64, INT64], b, [INT64, FLOAT64], p, [0], [0])
        assert_same_multiset(mj, hj)
