"""CPU suite: the oracle's multi-column join keys / GROUP BY lists against an independent Python restatement
(dict of key tuples) — the two-implementations cross-check of SURVEY §8c(iii) — plus the reference's own
`group by a, b` golden (executor/aggregate_test.go:64-66)."""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200.chunk import FLOAT64, INT64, UINT64, Column
from util import gen_col

INNER, LEFT, RIGHT = 0, 1, 2
COUNT, SUM, AVG, MAX, MIN, FIRSTROW = range(6)


def key_of(cols, types, idx, r):
    """(flag, raw bits) per key column as codec.encodeHashChunkRowIdx sees it (util/codec/codec.go:212-240); None if any NULL"""
    out = []
    for c in idx:
        if not cols[c].not_null()[r]:
            return None
        raw = int(cols[c].raw()[r])
        if types[c] == FLOAT64:
            out.append(("f", raw))
        elif types[c] == UINT64 and raw >> 63:
            out.append(("u", raw))
        else:
            out.append(("i", raw))
    return tuple(out)


def py_join(jt, bt, b, pt, p, bk, pk):
    """inner / probe-side-outer hash join on key tuples; rows as (probe cols ++ build cols) tuples with None for NULL"""
    table = {}
    for r in range(b[0].length):
        k = key_of(b, bt, bk, r)
        if k is not None:
            table.setdefault(k, []).append(r)
    val = lambda cols, r: tuple(None if not c.not_null()[r] else int(c.raw()[r]) for c in cols)
    out = []
    for r in range(p[0].length):
        k = key_of(p, pt, pk, r)
        m = table.get(k, []) if k is not None else []
        for br in m:
            out.append(val(p, r) + val(b, br))
        if not m and jt != INNER:
            out.append(val(p, r) + (None,) * len(b))
    return sorted(out, key=str)


@pytest.mark.parametrize("jt", [INNER, LEFT])
def test_oracle_multi_column_join_keys(jt):
    rng = np.random.default_rng(7 + jt)
    nb, npr = 700, 3000
    bt, pt = [INT64, UINT64, FLOAT64, INT64], [FLOAT64, INT64, INT64]
    b = [gen_col(rng, INT64, nb, 0.05, -4, 4), Column(UINT64, rng.integers(0, 6, nb).astype(np.uint64), rng.random(nb) > 0.05),
         Column(FLOAT64, rng.integers(0, 3, nb) * 0.5, rng.random(nb) > 0.05), Column(INT64, np.arange(nb))]
    p = [Column(FLOAT64, rng.integers(0, 4, npr) * 0.5, rng.random(npr) > 0.05), gen_col(rng, INT64, npr, 0.05, -5, 5),
         Column(INT64, rng.integers(-1, 7, npr), rng.random(npr) > 0.05)]
    bk, pk = [0, 1, 2], [1, 2, 0]
    # probe side is the left child (outer_is_right = False): output = probe cols ++ build cols
    got = O.hash_join(jt, False, bt, b, pt, p, bk, pk)
    rows = [tuple(None if v is None else int(np.array(v).astype(np.float64).view(np.uint64)) if isinstance(v, float) else
                  (int(v) & ((1 << 64) - 1)) for v in r) for r in got.rows()]
    assert sorted(rows, key=str) == py_join(jt, bt, b, pt, p, bk, pk)


def test_oracle_group_by_two_columns_reference_golden():
    # executor/aggregate_test.go:64-66: (a,b) rows with b>0; count(a) group by a, b -> sorted "1","1","1","3"
    rows = [(1, 1), (3, 3), (3, 2), (2, 1), (1, 1), (1, 1)]
    a, b = Column(INT64, [r[0] for r in rows]), Column(INT64, [r[1] for r in rows])
    for workers in (1, 3):
        rc, out = O.hash_agg([INT64, INT64], [a, b], [0, 1], [(COUNT, 0), (FIRSTROW, 0), (FIRSTROW, 1)], workers)
        assert rc == 0 and sorted(out.rows()) == [(1, 2, 1), (1, 3, 2), (1, 3, 3), (3, 1, 1)]
    # "... group by a, b order by a" -> "3","1","1","1" (:66)
    assert [r[0] for r in sorted(out.rows(), key=lambda r: (r[1], r[2]))] == [3, 1, 1, 1]


def test_oracle_group_by_multi_column_vs_python():
    rng = np.random.default_rng(12)
    n = 20000
    k1, k2 = gen_col(rng, INT64, n, 0.05, -5, 5), gen_col(rng, UINT64, n, 0.05, 0, 7)
    k3 = Column(FLOAT64, rng.integers(0, 4, n) * 0.5, rng.random(n) > 0.05)
    v = gen_col(rng, INT64, n, 0.1, -100, 100)
    types, cols = [INT64, UINT64, FLOAT64, INT64], [k1, k2, k3, v]
    funcs = [(FIRSTROW, 0), (FIRSTROW, 1), (FIRSTROW, 2), (COUNT, -1), (COUNT, 3), (SUM, 3), (MAX, 3), (MIN, 3)]
    rc, out = O.hash_agg(types, cols, [0, 1, 2], funcs, 3)
    assert rc == 0
    groups = {}
    kv = lambda c, r: None if not c.not_null()[r] else c.values[r].item()
    for r in range(n):
        g = groups.setdefault((kv(k1, r), kv(k2, r), kv(k3, r)), [0, 0, None, None, None])
        g[0] += 1
        x = kv(v, r)
        if x is not None:
            g[1] += 1
            g[2] = x if g[2] is None else g[2] + x
            g[3] = x if g[3] is None else max(g[3], x)
            g[4] = x if g[4] is None else min(g[4], x)
    want = sorted((k + tuple(s) for k, s in groups.items()), key=str)
    assert sorted(out.rows(), key=str) == want
