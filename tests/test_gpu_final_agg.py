"""GPU parity for FinalMode HashAggExec over pushed-down partial results (tq_agg_create_final):
the partial rows come from the oracle's restatement of the coprocessor (store/mockstore/mocktikv/aggregate.go), the GPU merges
them through the C-ABI, and the result must equal the oracle's FinalMode executor (aggfuncs/builder.go:50-62,86-109) — and,
by the planner's split identity (planner/core/task.go:564-625), the Complete-mode aggregation of the raw rows."""
import numpy as np
import pytest

import oracle_py as O
from test_oracle_final_agg import AVG, COUNT, FIRSTROW, MAX, MIN, SUM, assert_agg_equal, final_funcs, partial_rows
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, UINT64, Column
from tinysql_b200.executor import HashAggFinalExec, MockDataSource
from util import gen_col

pytestmark = pytest.mark.gpu


def run_final(ptypes, pcols, gb, ff, chunk=1024, est=0):
    a = HashAggFinalExec(MockDataSource(ptypes, pcols, chunk), gb, ff, est)
    a.Open()
    got = a.drain()
    a.Close()
    rc, want = O.hash_agg_final(ptypes, pcols, gb, ff)
    assert rc == 0
    return got, want


def test_avg_final_mode_reference_golden(lib):
    # TestAvgFinalMode (expression/aggregation/aggregation_test.go:86-111)
    i = np.arange(1, 101, dtype=np.int64)
    got, want = run_final([INT64, INT64], [Column(INT64, i), Column(INT64, i * i)], [], [(AVG, 0, 1)])
    assert got.rows() == want.rows() == [(67,)]
    cnt = Column(INT64, [2, 5, 0])
    s = Column(FLOAT64, [3.0, 0.0, 0.0], [True, False, False])
    got, want = run_final([INT64, FLOAT64], [cnt, s], [], [(AVG, 0, 1), (COUNT, 0, -1)])
    assert got.rows() == want.rows() == [(1.5, 7)]


@pytest.mark.parametrize("regions", [1, 5])
def test_final_of_partials_fixed_width(lib, regions):
    rng = np.random.default_rng(300 + regions)
    n = 200000
    k1 = gen_col(rng, INT64, n, 0.05, 0, 5000)
    k2 = gen_col(rng, UINT64, n, 0.05, 0, 3)
    xi = gen_col(rng, INT64, n, 0.2, -1000, 1000)
    xf = Column(FLOAT64, rng.integers(-1000, 1000, n).astype(np.float64) * 0.25, rng.random(n) > 0.2)
    types, cols = [INT64, UINT64, INT64, FLOAT64], [k1, k2, xi, xf]
    bounds = [0] + sorted(rng.integers(0, n, regions - 1).tolist()) + [n]
    for group_by in ([0], [0, 1], []):
        funcs = [(COUNT, -1), (COUNT, 2), (SUM, 2), (AVG, 2), (SUM, 3), (AVG, 3), (MAX, 2), (MIN, 3), (MAX, 1)] + [(FIRSTROW, g) for g in group_by]
        ptypes, part = partial_rows(types, cols, group_by, funcs, bounds)
        ff, gb = final_funcs(funcs, len(group_by))
        got, want = run_final(ptypes, part.cols, gb, ff, chunk=1000, est=5000)
        nf = len(funcs) - len(group_by)
        keys = list(range(nf, len(funcs)))
        assert_agg_equal(got, want, keys, approx_cols=(4, 5))
        rc, complete = O.hash_agg(types, cols, group_by, funcs)
        assert rc == 0
        assert_agg_equal(got, complete, keys, approx_cols=(4, 5))


def test_final_of_partials_string_and_float_columns(lib):
    # partial MAX / MIN / FIRSTROW columns keep the argument's own chunk layout (FLOAT slots, var-len cells); string GROUP BY items
    rng = np.random.default_rng(41)
    n = 50000
    g = Column(BYTES, [(b"g%03d" % v) if ok else None for v, ok in zip(rng.integers(0, 400, n), rng.random(n) > 0.05)])
    s = Column(BYTES, [(b"v%04d" % v) if ok else None for v, ok in zip(rng.integers(0, 3000, n), rng.random(n) > 0.2)])
    f = Column(FLOAT32, rng.integers(-50, 50, n).astype(np.float32) * np.float32(0.5), rng.random(n) > 0.2)
    k = gen_col(rng, INT64, n, 0.1, 0, 4)
    types, cols = [BYTES, BYTES, FLOAT32, INT64], [g, s, f, k]
    bounds = [0, 7, 20000, 20001, 45000, n]
    for group_by in ([0], [0, 3], []):
        funcs = [(COUNT, 1), (MAX, 1), (MIN, 1), (MAX, 2), (MIN, 2), (SUM, 2), (AVG, 2), (FIRSTROW, 1)] + [(FIRSTROW, c) for c in group_by]
        ptypes, part = partial_rows(types, cols, group_by, funcs, bounds)
        ff, gb = final_funcs(funcs, len(group_by))
        got, want = run_final(ptypes, part.cols, gb, ff, chunk=777)
        nf = len(funcs) - len(group_by)
        if group_by:
            # FIRSTROW of a non-key column returns *a* row's value (DESIGN §3 deviation ii): compare it for membership only
            keys = list(range(nf, len(funcs)))
            fr = 7
            gk, wk = {tuple(r[c] for c in keys): r for r in got.rows()}, {tuple(r[c] for c in keys): r for r in want.rows()}
            assert set(gk) == set(wk)
            for key, wr in wk.items():
                for i, (x, y) in enumerate(zip(gk[key], wr)):
                    if i == fr:
                        continue
                    if i in (5, 6) and x is not None and y is not None:
                        assert abs(x - y) <= 1e-9 * max(1.0, abs(y))
                    else:
                        assert x == y, (key, i, x, y)
        else:
            assert got.num_rows() == want.num_rows() == 1
            assert got.rows()[0][:5] == want.rows()[0][:5]


def test_final_mode_empty_input_and_errors(lib):
    got, want = run_final([INT64, INT64, FLOAT64], [Column(INT64, []), Column(INT64, []), Column(FLOAT64, [])], [],
                          [(COUNT, 0, -1), (AVG, 1, 2), (SUM, 2, -1)])
    assert got.rows() == want.rows() == [(0, None, None)]
    a = HashAggFinalExec(MockDataSource([FLOAT64], [Column(FLOAT64, [1.0])]), [], [(COUNT, 0, -1)])
    with pytest.raises(L.TQError) as ei:   # a partial COUNT is a BIGINT column
        a.Open()
    assert ei.value.status == L.TQ_ERR_UNSUPPORTED_TYPE
