"""CPU suite: the oracle's restatement of pushed-down partial aggregation (store/mockstore/mocktikv/aggregate.go) and of the
FinalMode HashAggExec that consumes its rows (aggfuncs/builder.go:50-62,86-109) — SURVEY §8 f4.  Pinned by the reference's
TestAvgFinalMode (expression/aggregation/aggregation_test.go:86-111) and by the identity the planner relies on when it splits an
aggregation (planner/core/task.go:564-625): Final(Partial1(region 1) ++ ... ++ Partial1(region R)) == Complete(all rows)."""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, UINT64, Chunk, Column
from util import gen_col

COUNT, SUM, AVG, MAX, MIN, FIRSTROW = range(6)


def final_funcs(funcs, n_group_by):
    """the FinalMode descriptors over the coprocessor's output schema: partial columns in function order, then GROUP BY columns"""
    out, c = [], 0
    for f, _ in funcs:
        if f == AVG:
            out.append((f, c, c + 1))
            c += 2
        else:
            out.append((f, c, -1))
            c += 1
    return out, list(range(c, c + n_group_by))


def slice_cols(cols, lo, hi):
    return [Column(c.tp, c.values[lo:hi], c.not_null()[lo:hi]) for c in cols]


def partial_rows(types, cols, group_by, funcs, bounds):
    """one coprocessor response per region [lo, hi): returns (partial schema types, concatenated partial rows)"""
    chunks, ptypes = [], None
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        rc, ptypes, part = O.cop_partial_agg(types, slice_cols(cols, lo, hi), group_by, funcs)
        assert rc == 0
        chunks.append(part)
    return ptypes, Chunk.concat(chunks, ptypes)


def by_key(chunk, key_cols):
    return {tuple(r[c] for c in key_cols): r for r in chunk.rows()}


def assert_agg_equal(got, want, key_cols, approx_cols=()):
    assert got.num_rows() == want.num_rows()
    assert [c.tp for c in got.cols] == [c.tp for c in want.cols]
    g, w = by_key(got, key_cols), by_key(want, key_cols)
    assert set(g) == set(w)
    for k, wr in w.items():
        for i, (x, y) in enumerate(zip(g[k], wr)):
            if i in approx_cols and x is not None and y is not None:
                assert abs(x - y) <= 1e-9 * max(1.0, abs(y)), (k, i, x, y)
            else:
                assert x == y, (k, i, x, y)


def test_avg_final_mode_reference_golden():
    # TestAvgFinalMode: partial rows (count = i, sum = i*i), i = 1..100 -> AVG = 338350 / 5050 = 67 (integer division)
    i = np.arange(1, 101, dtype=np.int64)
    rc, got = O.hash_agg_final([INT64, INT64], [Column(INT64, i), Column(INT64, i * i)], [], [(AVG, 0, 1)])
    assert rc == 0 and got.rows() == [(67,)]
    # a partial row with a NULL sum (an all-NULL group in that region) is skipped together with its count (func_avg.go:93-103)
    cnt = Column(INT64, [2, 5, 0])
    s = Column(FLOAT64, [3.0, 0.0, 0.0], [True, False, False])
    rc, got = O.hash_agg_final([INT64, FLOAT64], [cnt, s], [], [(AVG, 0, 1), (COUNT, 0, -1)])
    assert rc == 0 and got.rows() == [(1.5, 7)]


def test_cop_partial_layout():
    # aggregate.go:98-108: per group the GetPartialResult datums in function order (AVG: count, sum), then the GROUP BY values
    k = Column(INT64, [1, 2, 1, 0, 2, 1], [True, True, True, False, True, True])
    x = Column(FLOAT64, [1.0, 2.0, 3.0, 4.0, 0.0, 5.0], [True, True, True, True, False, True])
    rc, ptypes, part = O.cop_partial_agg([INT64, FLOAT64], [k, x], [0], [(COUNT, 1), (AVG, 1), (SUM, 1), (MAX, 1), (FIRSTROW, 0)])
    assert rc == 0 and ptypes == [INT64, INT64, FLOAT64, FLOAT64, FLOAT64, INT64, INT64]
    assert part.rows() == [(3, 3, 9.0, 9.0, 5.0, 1, 1), (1, 1, 2.0, 2.0, 2.0, 2, 2), (1, 1, 4.0, 4.0, 4.0, None, None)]
    # a group whose argument is NULL throughout: COUNT 0, AVG (0, NULL), SUM NULL
    rc, _, part = O.cop_partial_agg([INT64, FLOAT64], [Column(INT64, [7]), Column(FLOAT64, [0.0], [False])], [0], [(COUNT, 1), (AVG, 1), (SUM, 1)])
    assert part.rows() == [(0, 0, None, None, 7)]


@pytest.mark.parametrize("regions", [1, 3, 17])
def test_final_of_partials_equals_complete(regions):
    rng = np.random.default_rng(100 + regions)
    n = 20000
    k1 = gen_col(rng, INT64, n, 0.05, 0, 300)
    k2 = gen_col(rng, UINT64, n, 0.05, 0, 3)
    xi = gen_col(rng, INT64, n, 0.2, -1000, 1000)
    xf = Column(FLOAT64, rng.integers(-1000, 1000, n).astype(np.float64) * 0.25, rng.random(n) > 0.2)   # sums are exact in any order
    f32 = Column(FLOAT32, rng.integers(-50, 50, n).astype(np.float32) * np.float32(0.5), rng.random(n) > 0.1)
    s = Column(BYTES, [(b"w%03d" % v) if ok else None for v, ok in zip(rng.integers(0, 500, n), rng.random(n) > 0.1)])
    types, cols = [INT64, UINT64, INT64, FLOAT64, FLOAT32, BYTES], [k1, k2, xi, xf, f32, s]
    bounds = [0] + sorted(rng.integers(0, n, regions - 1).tolist()) + [n]
    for group_by in ([0], [0, 1], [5], []):
        funcs = [(COUNT, -1), (COUNT, 2), (SUM, 2), (AVG, 2), (SUM, 3), (AVG, 3), (MAX, 2), (MIN, 3), (MAX, 4), (MIN, 5), (MAX, 5), (AVG, 4)]
        funcs += [(FIRSTROW, g) for g in group_by]
        ptypes, part = partial_rows(types, cols, group_by, funcs, bounds)
        ff, gb = final_funcs(funcs, len(group_by))
        rc, got = O.hash_agg_final(ptypes, part.cols, gb, ff)
        assert rc == 0
        rc, want = O.hash_agg(types, cols, group_by, funcs)
        assert rc == 0
        nf = len(funcs) - len(group_by)
        assert_agg_equal(got, want, list(range(nf, len(funcs))), approx_cols=(4, 5, 11))


def test_final_mode_empty_input_default_row():
    rc, got = O.hash_agg_final([INT64, INT64, FLOAT64], [Column(INT64, []), Column(INT64, []), Column(FLOAT64, [])], [],
                               [(COUNT, 0, -1), (AVG, 1, 2), (SUM, 2, -1)])
    assert rc == 0 and got.rows() == [(0, None, None)]
