"""CPU suite: property tests (hypothesis) of the oracle — the invariants the GPU suite relies on at sizes where no second
implementation is run: build-side choice is result-neutral for inner joins, an outer join is the inner join plus exactly
the unmatched outer rows, partial/final splitting is result-neutral, COUNT sums to the non-NULL rows."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle_py as O
from tinysql_b200.chunk import FLOAT64, INT64, UINT64, Column

INNER, LEFT, RIGHT = 0, 1, 2
COUNT, SUM, AVG, MAX, MIN, FIRSTROW = range(6)

cell = st.one_of(st.none(), st.integers(-3, 6))


def table(rows, ncols):
    return [Column(INT64, [0 if r[c] is None else r[c] for r in rows], [r[c] is not None for r in rows]) for c in range(ncols)]


rows2 = st.lists(st.tuples(cell, cell), min_size=0, max_size=40)


def key(r):
    return tuple((0, 0) if v is None else (1, v) for v in r)


@settings(max_examples=150, deadline=None)
@given(rows2, rows2, st.integers(0, 1), st.integers(0, 1))
def test_inner_join_is_build_side_neutral_and_outer_adds_exactly_the_misses(lhs, rhs, lk, rk):
    l, r = table(lhs, 2), table(rhs, 2)
    t2 = [INT64, INT64]
    # build on the right child (probe = left) vs build on the left child (probe = right): same rows lhs ++ rhs
    a = O.hash_join(INNER, False, t2, r, t2, l, [rk], [lk]).rows()
    b = O.hash_join(INNER, True, t2, l, t2, r, [lk], [rk]).rows()
    assert sorted(a, key=key) == sorted(b, key=key)
    # reference definition: pairs with equal non-NULL keys
    want = [x + y for x in map(tuple, lhs) for y in map(tuple, rhs) if x[lk] is not None and x[lk] == y[rk]]
    assert sorted(a, key=key) == sorted(want, key=key)
    # left outer = inner + (unmatched left rows ++ NULLs), each exactly once
    lo = O.hash_join(LEFT, False, t2, r, t2, l, [rk], [lk]).rows()
    rkeys = {y[rk] for y in rhs if y[rk] is not None}
    misses = [tuple(x) + (None, None) for x in lhs if x[lk] is None or x[lk] not in rkeys]
    assert sorted(lo, key=key) == sorted(want + misses, key=key)
    # right outer (outer = right child, build = left child)
    ro = O.hash_join(RIGHT, True, t2, l, t2, r, [lk], [rk]).rows()
    lkeys = {x[lk] for x in lhs if x[lk] is not None}
    rmiss = [(None, None) + tuple(y) for y in rhs if y[rk] is None or y[rk] not in lkeys]
    assert sorted(ro, key=key) == sorted(want + rmiss, key=key)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(cell, cell, st.one_of(st.none(), st.integers(-50, 50))), min_size=0, max_size=60), st.integers(1, 5), st.booleans())
def test_group_by_matches_a_dictionary_and_partial_split_is_neutral(rows, workers, two_keys):
    cols = table(rows, 3)
    gb = [0, 1] if two_keys else [0]
    funcs = [(COUNT, -1), (COUNT, 2), (SUM, 2), (MAX, 2), (MIN, 2), (AVG, 2)] + [(FIRSTROW, g) for g in gb]
    rc1, one = O.hash_agg([INT64] * 3, cols, gb, funcs, 1)
    rc2, many = O.hash_agg([INT64] * 3, cols, gb, funcs, workers)
    assert rc1 == rc2 == 0
    assert sorted(one.rows(), key=key) == sorted(many.rows(), key=key)
    groups = {}
    for r in rows:
        g = groups.setdefault(tuple(r[c] for c in gb), [0, []])
        g[0] += 1
        if r[2] is not None:
            g[1].append(r[2])
    want = []
    for k, (n, vals) in groups.items():
        s = sum(vals) if vals else None
        avg = None if not vals else int(abs(s) // len(vals)) * (1 if s >= 0 else -1)  # Go truncating division (func_avg.go:53)
        want.append((n, len(vals), s, max(vals) if vals else None, min(vals) if vals else None, avg) + k)
    assert sorted(one.rows(), key=key) == sorted(want, key=key)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.tuples(st.integers(0, (1 << 64) - 1), st.integers(-(1 << 63), (1 << 63) - 1)), min_size=1, max_size=30))
def test_signed_unsigned_key_equality_is_numeric(pairs):
    """util/codec/codec.go:219-231: an unsigned and a signed key are equal iff they denote the same number"""
    u = Column(UINT64, np.array([p[0] for p in pairs], dtype=np.uint64))
    i = Column(INT64, np.array([p[1] for p in pairs], dtype=np.int64))
    got = O.hash_join(INNER, False, [UINT64], [u], [INT64], [i], [0], [0]).num_rows()
    want = sum(1 for a in pairs for b in pairs if b[1] >= 0 and a[0] == b[1])
    assert got == want


cond_st = st.tuples(st.integers(0, 5), st.integers(0, 3), st.one_of(st.integers(0, 3), st.none()), st.integers(-2, 5))


@settings(max_examples=150, deadline=None)
@given(rows2, rows2, st.lists(cond_st, min_size=1, max_size=2), st.integers(0, 2))
def test_other_conditions_filter_matches_and_turn_all_failed_outer_rows_into_misses(lhs, rhs, conds, jt):
    """joiner.go:155-167,225-248: OtherConditions filter the joined rows of one outer row; an outer row whose joined rows
    all fail is emitted once with NULLs (outer joins) or not at all (inner join)"""
    l, r = table(lhs, 2), table(rhs, 2)
    t2 = [INT64, INT64]
    oc = [(op, a, b) if b is not None else (op, a, None, INT64, k) for op, a, b, k in conds]
    if jt == RIGHT:
        got = O.hash_join(RIGHT, True, t2, l, t2, r, [0], [0], None, oc).rows()
    else:
        got = O.hash_join(jt, False, t2, r, t2, l, [0], [0], None, oc).rows()

    def ok(row):
        for op, a, b, k in conds:
            x, y = row[a], (row[b] if b is not None else k)
            if x is None or y is None:
                return False
            if not [x < y, x <= y, x > y, x >= y, x == y, x != y][op]:
                return False
        return True
    want = []
    outer, inner = (rhs, lhs) if jt == RIGHT else (lhs, rhs)
    for o in map(tuple, outer):
        rows = []
        for i in map(tuple, inner):
            if o[0] is not None and o[0] == i[0]:
                joined = (i + o) if jt == RIGHT else (o + i)
                if ok(joined):
                    rows.append(joined)
        if not rows and jt != INNER:
            rows = [((None, None) + o) if jt == RIGHT else (o + (None, None))]
        want += rows
    assert sorted(got, key=key) == sorted(want, key=key)


def test_other_condition_reference_golden():
    # executor/join_test.go:137-139: t1 join t on t.a = t1.a and t.a < t1.b -> "1 2 1 1","1 3 1 1","1 4 1 1","3 4 3 3"
    t1 = [(1, 2), (1, 3), (1, 4), (3, 4), (4, 5)]
    t = [(1, 1), (2, 2), (3, 3)]
    l, r = table(t1, 2), table(t, 2)
    got = O.hash_join(INNER, False, [INT64, INT64], r, [INT64, INT64], l, [0], [0], None, [(0, 2, 1)]).rows()  # t.a (col 2) < t1.b (col 1)
    assert sorted(got) == [(1, 2, 1, 1), (1, 3, 1, 1), (1, 4, 1, 1), (3, 4, 3, 3)]


# ---------------------------------------------------------------------------------------------- sort / top-n / merge join
by_item = st.tuples(st.integers(0, 1), st.booleans())


@settings(max_examples=150, deadline=None)
@given(rows2, st.lists(by_item, min_size=0, max_size=3), st.integers(0, 6), st.integers(-1, 8))
def test_sort_is_a_stable_permutation_ordered_by_the_by_items(rows, by, off, cnt):
    t2 = [INT64, INT64]
    cols = table(rows, 2) + [Column(INT64, np.arange(len(rows)))]            # row id: makes the permutation visible
    got = O.sort(t2 + [INT64], cols, by).rows()
    assert sorted(r[2] for r in got) == list(range(len(rows)))                # a permutation of the child rows
    assert [r[:2] for r in got] == [tuple(rows[r[2]]) for r in got]           # rows travel whole

    def cmp_key(r):   # cmpNull + the value, sign flipped for Desc (sort.go:115-129)
        return tuple(((0, 0) if r[c] is None else (1, r[c])) for c, _ in by)
    for a, b in zip(got, got[1:]):
        for c, desc in by:
            ka, kb = ((0, 0) if a[c] is None else (1, a[c])), ((0, 0) if b[c] is None else (1, b[c]))
            if ka != kb:
                assert (ka > kb) if desc else (ka < kb)
                break
        else:
            assert a[2] < b[2]                                                 # equal keys keep child order
    # sorting the sorted rows again changes nothing (idempotence), TopN is a window of the order
    again = O.sort(t2 + [INT64], [Column(INT64, [0 if r[c] is None else r[c] for r in got], [r[c] is not None for r in got]) for c in range(2)] +
                   [Column(INT64, [r[2] for r in got])], by).rows()
    assert again == got
    lo = min(off, len(got))
    assert O.sort(t2 + [INT64], cols, by, off, cnt).rows() == (got[lo:] if cnt < 0 else got[lo: lo + cnt])


@settings(max_examples=150, deadline=None)
@given(rows2, rows2, st.sampled_from([(INNER, False), (INNER, True), (LEFT, False), (RIGHT, True)]), st.lists(st.booleans(), min_size=40, max_size=40))
def test_merge_join_of_sorted_children_is_the_hash_join_in_outer_order(inner_rows, outer_rows, mode, sel_bits):
    jt, oir = mode
    t2 = [INT64, INT64]
    inner = O.sort(t2, table(inner_rows, 2), [(0, False)])
    outer = O.sort(t2, table(outer_rows, 2), [(0, False)])
    sel = np.array(sel_bits[: len(outer_rows)], dtype=np.uint8)
    mj = O.merge_join(jt, oir, t2, inner.cols, t2, outer.cols, [0], [0], sel).rows()
    hj = O.hash_join(jt, oir, t2, inner.cols, t2, outer.cols, [0], [0], sel).rows()
    assert sorted(mj, key=key) == sorted(hj, key=key)                         # the same rows as the hash join ...
    o_cols = slice(2, 4) if oir else slice(0, 2)
    orows = outer.rows()
    pos, it = [], iter(range(len(orows)))
    cur = next(it, None)
    for r in mj:                                                               # ... in the order of the outer child
        while cur is not None and tuple(orows[cur]) != r[o_cols]:
            cur = next(it, None)
        assert cur is not None, "an output row does not follow the outer child's order"
        pos.append(cur)
    assert pos == sorted(pos)
