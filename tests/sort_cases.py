"""Known-answer cases for SortExec / TopNExec / MergeJoinExec transcribed from the reference's own tests (inputs as the tables the
SQL builds, expected rows as testkit.Rows prints them).  Run against the oracle (test_oracle_sort_merge.py, CPU) and against the
GPU operators (test_gpu_sort_merge.py)."""
from tinysql_b200.chunk import BYTES, INT64, UINT64, Column

INNER, LEFT, RIGHT = 0, 1, 2


def icol(vals):
    return Column(INT64, [0 if v is None else v for v in vals], [v is not None for v in vals])


def _order_table():
    # executor_test.go:505-521 — fillData (1 hello, 2 hello) + the inserts of "Test limit + order by"
    ids = [1, 2] + list(range(3, 11)) + [10086] + list(range(11, 21)) + list(range(21, 31)) + [1501]
    names = [b"hello", b"hello"] + [b"zz"] * 8 + [b"hi"] + [b"hh"] * 10 + [b"zz"] * 10 + [b"aa"]
    return [INT64, BYTES], [Column(INT64, ids), Column(BYTES, names)]


# (name, types, cols, by [(col, desc)], offset, count, expected rows)
SORT_CASES = [
    # TestSelectOrderBy (executor_test.go:472-523)
    ("order by id desc limit 1", [INT64, BYTES], [Column(INT64, [1, 2]), Column(BYTES, [b"hello", b"hello"])], [(0, True)], 0, 1, [(2, b"hello")]),
    ("order by name, id limit 1 offset 0", [INT64, BYTES], [Column(INT64, [1, 2]), Column(BYTES, [b"hello", b"hello"])], [(1, False), (0, False)], 0, 1, [(1, b"hello")]),
    ("limit overflow: limit 100 offset 0", [INT64, BYTES], [Column(INT64, [2, 1]), Column(BYTES, [b"hello", b"hello"])], [(1, False), (0, False)], 0, 100,
     [(1, b"hello"), (2, b"hello")]),
    ("offset overflow: limit 1 offset 100", [INT64, BYTES], [Column(INT64, [1, 2]), Column(BYTES, [b"hello", b"hello"])], [(1, False), (0, False)], 100, 1, []),
    ("limit 18446744073709551615", [INT64, BYTES], [Column(INT64, [1, 2]), Column(BYTES, [b"hello", b"hello"])], [(1, False), (0, False)], 0, (1 << 63) - 1,
     [(1, b"hello"), (2, b"hello")]),
    ("order by name, id limit 1 offset 3", *_order_table(), [(1, False), (0, False)], 3, 1, [(11, b"hh")]),
    # executor_test.go:548-549
    ("order by b", [INT64, INT64], [icol([2, 1]), icol([2, 1])], [(1, False)], 0, -1, [(1, 1), (2, 2)]),
    ("order by a desc", [INT64, INT64], [icol([1, 2]), icol([1, 2])], [(0, True)], 0, -1, [(2, 2), (1, 1)]),
    # NULLs sort first ascending and last descending (util/chunk/compare.go:45-53, sort.go:120-122)
    ("nulls first", [INT64], [icol([3, None, 1, None, 2])], [(0, False)], 0, -1, [(None,), (None,), (1,), (2,), (3,)]),
    ("nulls last when desc", [INT64], [icol([3, None, 1, None, 2])], [(0, True)], 0, -1, [(3,), (2,), (1,), (None,), (None,)]),
]


def _t(rows, ncols=2):
    return [icol([r[c] for r in rows]) for c in range(ncols)]


# (name, join_type, outer_is_right, inner types/cols, outer types/cols, inner_keys, outer_keys, selected or None, expected rows = left ++ right)
MERGE_CASES = [
    # TestMergeJoin (merge_join_test.go:238-322): t = (1,1),(2,2); t1 = (2,3),(4,4)
    ("t left outer join t1 on t.c1 = t1.c1", LEFT, False, [INT64, INT64], _t([(2, 3), (4, 4)]), [INT64, INT64], _t([(1, 1), (2, 2)]), [0], [0], None,
     [(1, 1, None, None), (2, 2, 2, 3)]),
    ("t1 right outer join t on t.c1 = t1.c1", RIGHT, True, [INT64, INT64], _t([(2, 3), (4, 4)]), [INT64, INT64], _t([(1, 1), (2, 2)]), [0], [0], None,
     [(None, None, 1, 1), (2, 3, 2, 2)]),
    ("t right outer join t1 on t.c1 = t1.c1", RIGHT, True, [INT64, INT64], _t([(1, 1), (2, 2)]), [INT64, INT64], _t([(2, 3), (4, 4)]), [0], [0], None,
     [(2, 2, 2, 3), (None, None, 4, 4)]),
    # ... left outer join t1 on t.c1 = t1.c1 and t.c1 != 1: the outer-side condition is the outer filter (merge_join.go:262)
    ("left outer join with outer filter t.c1 != 1", LEFT, False, [INT64, INT64], _t([(2, 3), (4, 4)]), [INT64, INT64], _t([(1, 1), (2, 2)]), [0], [0], [0, 1],
     [(1, 1, None, None), (2, 2, 2, 3)]),
    # t1 (c1 int): (1),(1),(1) self join -> 9 rows "1 1"
    ("self join of three equal keys", INNER, False, [INT64], _t([(1,), (1,), (1,)], 1), [INT64], _t([(1,), (1,), (1,)], 1), [0], [0], None, [(1, 1)] * 9),
    # t(c1 int) = (1), t1(c1 int unsigned) = (1): mixed signedness compares by value (builtin_compare.go:541-560)
    ("int joined with int unsigned", INNER, False, [UINT64], [Column(UINT64, [1])], [INT64], [Column(INT64, [1])], [0], [0], None, [(1, 1)]),
    # t(a, b) = (1, 2): t right join t t1 on t.a = t1.b -> "<nil> 2" (columns t.a, t1.b)
    ("right join with no match", RIGHT, True, [INT64], [icol([1])], [INT64], [icol([2])], [0], [0], None, [(None, 2)]),
    # t(a, b) pk(a, b) = (1,1),(1,2),(1,3),(1,4); s(a) = (1): count(*) of t join s on t.a = s.a -> 4
    ("four outer rows share one inner key", INNER, False, [INT64], [icol([1])], [INT64, INT64], _t([(1, 1), (1, 2), (1, 3), (1, 4)]), [0], [0], None,
     [(1, 1, 1), (1, 2, 1), (1, 3, 1), (1, 4, 1)]),
    # TestMergejoinOrder (join_test.go:355-390): t1 = (1..5, 100), t2 = (100..500, 10000), left cond t1.a != 3: every row a miss, in order
    ("TestMergejoinOrder left outer", LEFT, False, [INT64, INT64], _t([(100 * i, 10000) for i in range(1, 6)]), [INT64, INT64], _t([(i, 100) for i in range(1, 6)]),
     [0], [0], [1, 1, 0, 1, 1], [(i, 100, None, None) for i in range(1, 6)]),
    # t(a, b) idx(a, b) = (1,1),(1,2),(2,1),(2,2) joined with itself on b and a
    ("two key columns", INNER, False, [INT64, INT64], _t([(1, 1), (1, 2), (2, 1), (2, 2)]), [INT64, INT64], _t([(1, 1), (1, 2), (2, 1), (2, 2)]), [0, 1], [0, 1], None,
     [(1, 1, 1, 1), (1, 2, 1, 2), (2, 1, 2, 1), (2, 2, 2, 2)]),
]
