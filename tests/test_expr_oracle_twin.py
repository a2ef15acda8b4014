"""CPU suite, part 3: the oracle's vectorized builtins against an independent restatement of the ROW versions
(expression/builtin_arithmetic.go, builtin_compare.go, builtin_op.go, builtin_control.go, builtin_other.go) written with
Python big integers — the reference's own differential strategy (expression/bench_test.go:415-504: vecEval* must equal
the row-at-a-time eval* on every non-NULL row and produce the same NULL mask), on the reference's operand ranges."""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200.chunk import FLOAT64, INT64, UINT64, Column
from util import gen_col

I64_MIN, I64_MAX, U64_MAX = -(1 << 63), (1 << 63) - 1, (1 << 64) - 1


def ival(col, i):
    v = int(col.values[i])
    return v


def row_arith_int(op, a, b, i):
    """returns (value, error) for one non-NULL row; mathematically exact range rules of the row versions"""
    x, y = ival(a, i), ival(b, i)
    ua, ub = a.tp == UINT64, b.tp == UINT64
    if op == 2 and (ua or ub):
        x, y = x % (1 << 64), y % (1 << 64)  # MultiplyIntUnsigned reads both operands as uint64 (builtin_arithmetic.go:397-414)
    r = x + y if op == 0 else (x - y if op == 1 else x * y)
    unsigned_result = ua or ub
    lo, hi = (0, U64_MAX) if unsigned_result else (I64_MIN, I64_MAX)
    if r < lo or r > hi:
        return None, True
    return r, False


@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("ta,tb", [(INT64, INT64), (UINT64, UINT64), (UINT64, INT64), (INT64, UINT64)])
def test_arith_int_matches_row_version(op, ta, tb):
    rng = np.random.default_rng(op * 17 + ta * 5 + tb)
    n = 1024
    # builtin_arithmetic_vec_test.go:47-76: operands in [MinInt64/2, MaxInt64/2] (signed) / [0, MaxInt64] (unsigned)
    lim = (1 << 31) if op == 2 else (1 << 62)
    mixed = (ta == UINT64) != (tb == UINT64)
    # signed/unsigned mixes use non-negative signed operands, like builtin_arithmetic_vec_test.go:53-76
    a = gen_col(rng, ta, n, 0.2, 0 if (ta == UINT64 or mixed) else -lim, lim)
    b = gen_col(rng, tb, n, 0.2, 0 if (tb == UINT64 or mixed) else -lim, lim)
    if op == 1 and (ta == UINT64 or tb == UINT64):
        # keep unsigned subtraction non-negative like the reference's generators do
        if ta == UINT64 and tb == UINT64:
            hi = np.maximum(a.values, b.values); lo = np.minimum(a.values, b.values)
            a, b = Column(UINT64, hi, a.not_null()), Column(UINT64, lo, b.not_null())
        elif ta == UINT64:
            b = Column(INT64, -np.abs(b.values), b.not_null())
        else:
            a = Column(INT64, np.abs(a.values) + (1 << 62), a.not_null()); b = Column(UINT64, b.values % np.uint64(1 << 61), b.not_null())
    rc, out = O.vec_arith_int(op, a, b)
    assert rc == 0
    nn = out.not_null()
    assert np.array_equal(nn, a.not_null() & b.not_null())
    for i in np.nonzero(nn)[0][:400]:
        want, err = row_arith_int(op, a, b, i)
        assert not err
        got = int(out.values[i])
        assert got == want, (i, got, want)


def test_arith_int_overflow_matches_row_version():
    cases = [(0, INT64, INT64, I64_MAX, 1), (0, INT64, INT64, I64_MIN, -1), (1, INT64, INT64, I64_MIN, 1), (2, INT64, INT64, 1 << 32, 1 << 32),
             (0, UINT64, UINT64, U64_MAX, 1), (1, UINT64, UINT64, 1, 2), (2, UINT64, UINT64, 1 << 33, 1 << 33), (0, UINT64, INT64, 3, -5),
             (0, INT64, UINT64, -5, 3), (1, UINT64, INT64, 1, 2), (0, INT64, INT64, 5, 6), (1, UINT64, INT64, 5, -6), (2, INT64, INT64, -3, 4),
             (2, UINT64, INT64, 3, -1), (2, INT64, UINT64, 0, 7), (2, UINT64, INT64, 6, 7)]
    for op, ta, tb, x, y in cases:
        a = Column(ta, np.array([x], dtype=np.uint64 if ta == UINT64 else np.int64))
        b = Column(tb, np.array([y], dtype=np.uint64 if tb == UINT64 else np.int64))
        want, err = row_arith_int(op, a, b, 0)
        rc, out = O.vec_arith_int(op, a, b)
        assert (rc != 0) == err, (op, ta, tb, x, y, rc)
        if not err:
            assert int(out.values[0]) == want


@pytest.mark.parametrize("ta,tb", [(INT64, INT64), (UINT64, UINT64), (UINT64, INT64), (INT64, UINT64)])
def test_compare_int_matches_row_version(ta, tb):
    # builtin_compare.go CompareInt (:525+): numeric comparison across signedness
    rng = np.random.default_rng(ta * 3 + tb)
    n = 2048
    a, b = gen_col(rng, ta, n), gen_col(rng, tb, n)
    b.values[:200] = a.values[:200].astype(b.values.dtype)
    for op, fn in enumerate([lambda x, y: x < y, lambda x, y: x <= y, lambda x, y: x > y, lambda x, y: x >= y, lambda x, y: x == y,
                             lambda x, y: x != y]):
        rc, out = O.vec_compare_int(op, a, b)
        nn = out.not_null()
        assert np.array_equal(nn, a.not_null() & b.not_null())
        for i in np.nonzero(nn)[0][:300]:
            assert int(out.values[i]) == int(fn(ival(a, i), ival(b, i)))


def test_real_ops_match_numpy():
    rng = np.random.default_rng(1)
    n = 4096
    a, b = gen_col(rng, FLOAT64, n), gen_col(rng, FLOAT64, n)
    b.values[:5] = 0.0
    nn = a.not_null() & b.not_null()
    for op, fn in enumerate([np.add, np.subtract, np.multiply]):
        rc, out, dz = O.vec_arith_real(op, a, b)
        assert rc == 0 and dz == 0 and np.array_equal(out.not_null(), nn)
        assert np.array_equal(out.values[nn], fn(a.values, b.values)[nn])
    rc, out, dz = O.vec_arith_real(3, a, b)  # division by zero -> NULL + warning (builtin_arithmetic_vec.go:369-375)
    exp_nn = nn & (b.values != 0)
    assert rc == 0 and np.array_equal(out.not_null(), exp_nn) and dz == int((nn & (b.values == 0)).sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        assert np.array_equal(out.values[exp_nn], (a.values / b.values)[exp_nn])
    assert O.vec_arith_real(2, Column(FLOAT64, [1e200]), Column(FLOAT64, [1e200]))[0] == 5  # ErrOverflow DOUBLE
    for op, fn in enumerate([np.less, np.less_equal, np.greater, np.greater_equal, np.equal, np.not_equal]):
        rc, out = O.vec_compare_real(op, a, b)
        assert np.array_equal(out.values[nn].astype(bool), fn(a.values, b.values)[nn])


def test_three_valued_logic_truth_tables():
    # MySQL 3-valued AND / OR (builtin_op.go): NULL AND 0 = 0, NULL AND 1 = NULL, NULL OR 1 = 1, NULL OR 0 = NULL
    vals = [(0, True), (1, True), (5, True), (0, False)]
    a = Column(INT64, [x[0] for x in vals for _ in vals], [x[1] for x in vals for _ in vals])
    b = Column(INT64, [y[0] for _ in vals for y in vals], [y[1] for _ in vals for y in vals])
    t = lambda v, nn: None if not nn else bool(v)
    for i, (op, name) in enumerate([(0, "and"), (1, "or")]):
        rc, out = O.vec_logic(op, a, b)
        res = out.tolist()
        for k in range(a.length):
            x, y = t(a.values[k], a.not_null()[k]), t(b.values[k], b.not_null()[k])
            if name == "and":
                want = 0 if (x is False or y is False) else (None if (x is None or y is None) else 1)
            else:
                want = 1 if (x is True or y is True) else (None if (x is None or y is None) else 0)
            assert res[k] == want, (name, x, y, res[k])


def test_control_other_and_filter():
    a = Column(INT64, [1, 0, 7, 3], [True, True, False, True])
    b = Column(INT64, [10, 20, 30, 40], [True, False, True, True])
    c = Column(INT64, [5, 6, 7, 8], [True, True, True, False])
    assert O.vec_if(a, b, c)[1].tolist() == [10, 6, 7, 40]          # IF(cond, b, c): NULL cond -> else branch
    assert O.vec_ifnull(b, c)[1].tolist() == [10, 6, 30, 40]
    assert O.vec_unary(4, a)[1].tolist() == [0, 0, 1, 0]            # ISNULL is never NULL
    assert O.vec_unary(0, a)[1].tolist() == [0, 1, None, 0]         # NOT
    assert O.vec_unary(2, a)[1].tolist() == [-1, 0, None, -3]       # unary minus
    assert O.vec_unary(2, Column(INT64, [I64_MIN]))[0] == 3          # -MinInt64 overflows (builtin_op_vec.go:236)
    # IN: found -> 1; not found with a NULL candidate -> NULL; else 0 (builtin_other_vec_generated.go:42-94)
    x = Column(INT64, [1, 2, 3, 4], [True, True, True, False])
    l1 = Column(INT64, [1, 9, 9, 9], [True, True, False, True])
    l2 = Column(UINT64, np.array([7, 2, 8, 4], dtype=np.uint64))
    assert O.vec_in_int(x, [l1, l2])[1].tolist() == [1, 1, None, None]
    neg = Column(INT64, [-1])
    big = Column(UINT64, np.array([U64_MAX], dtype=np.uint64))
    assert O.vec_in_int(neg, [big])[1].tolist() == [0]               # int64(-1) is not uint64(2^64-1)
    assert list(O.vec_filter_int(a)) == [1, 0, 0, 1]
