"""fused Selection + Projection (tq_expr_eval) against the oracle's per-builtin restatements composed the
way the reference composes them: VecEvalBool narrows the evaluation set filter by filter (expression/expression.go:205-279),
the projection sees only selected rows (executor/executor.go SelectionExec → ProjectionExec).  Bit-exact, same error kinds."""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200 import _lib as L
from tinysql_b200 import expression as E
from tinysql_b200.chunk import FLOAT64, INT64, UINT64, Column
from tinysql_b200.expression import Col, Const, ExprProgram, Func
from util import gen_col

pytestmark = pytest.mark.gpu

_TP = {"int": INT64, "uint": UINT64, "real": FLOAT64}


def _take(col, rows):
    return Column(col.tp, col.values[rows].copy(), col.not_null()[rows].copy())


class OracleError(Exception):
    def __init__(self, status):
        self.status = status


def _oracle(e, cols, rows, warn):
    """evaluates expression e over the rows `rows` of the chunk with the oracle's builtins -> Column of len(rows)"""
    n = len(rows)
    if isinstance(e, Col):
        return _take(cols[e.idx], rows)
    if isinstance(e, Const):
        dt = {"int": np.int64, "uint": np.uint64, "real": np.float64}[e.tp]
        if e.value is None:
            return Column(_TP[e.tp], np.zeros(n, dtype=dt), np.zeros(n, dtype=bool))
        return Column(_TP[e.tp], np.full(n, e.value, dtype=dt), np.ones(n, dtype=bool))
    a = [_oracle(x, cols, rows, warn) for x in e.args] if e.name != "in" else None
    real = any(x.tp == "real" for x in e.args)
    nm = e.name
    if nm in E._CMP:
        rc, out = (O.vec_compare_real if real else O.vec_compare_int)(E._CMP[nm], a[0], a[1])
    elif nm in E._ARITH:
        if real:
            rc, out, dz = O.vec_arith_real(E._ARITH[nm], a[0], a[1])
            warn[0] += dz
        else:
            rc, out = O.vec_arith_int(E._ARITH[nm], a[0], a[1])
    elif nm in ("and", "or"):
        rc, out = O.vec_logic(E.AND if nm == "and" else E.OR, a[0], a[1])
    elif nm == "not":
        rc, out = O.vec_unary(E.NOT_REAL if real else E.NOT_INT, a[0])
    elif nm == "neg":
        rc, out = O.vec_unary(E.MINUS_REAL if real else E.MINUS_INT, a[0])
    elif nm == "isnull":
        rc, out = O.vec_unary(E.ISNULL, a[0])
    elif nm == "if":
        rc, out = O.vec_if(a[0], a[1], a[2])
    elif nm == "ifnull":
        rc, out = O.vec_ifnull(a[0], a[1])
    else:  # in: the oracle's own builtinIn{Int,Real}Sig restatement, NOT the EQ/OR lowering the product uses
        a0 = _oracle(e.args[0], cols, rows, warn)
        lst = [_oracle(x, cols, rows, warn) for x in e.args[1:]]
        rc, out = (O.vec_in_real if real else O.vec_in_int)(a0, lst)
    if rc != 0:
        raise OracleError(rc)
    return out


def oracle_select_project(cols, filters, projections):
    n = cols[0].length
    rows = np.arange(n)
    nulls = np.zeros(n, dtype=bool)
    warn = [0]
    for f in filters:  # VecEvalBool
        v = _oracle(f, cols, rows, warn)
        nn = v.not_null()
        if f.tp == "real":
            zero = np.abs(v.values) < 0.5
            keep = nn & ~zero
        else:
            zero = v.values == 0
            keep = ~nn | ~zero
            nulls[rows[~nn]] = True
        rows = rows[keep]
    selected = np.zeros(n, dtype=np.uint8)
    selected[rows[~nulls[rows]]] = 1
    rows = rows[~nulls[rows]] if filters else rows
    outs = [_oracle(p, cols, rows, warn) for p in projections]
    return outs, selected, rows, warn[0]


def check(cols, filters, projections):
    prog = ExprProgram(len(cols), filters, projections)
    try:
        want, wsel, rows, wwarn = oracle_select_project(cols, filters, projections)
    except OracleError as oe:
        with pytest.raises(L.TQError) as ei:
            prog.run(cols)
        assert ei.value.status == oe.status
        return None
    got, sel, warn = prog.run(cols)
    if filters:
        assert np.array_equal(sel, wsel)
    assert warn == wwarn
    for g, w in zip(got, want):
        assert g.tp == w.tp
        gn = g.not_null()[rows]
        assert np.array_equal(gn, w.not_null())
        assert np.array_equal(g.values[rows][gn].view(np.uint64), w.values[gn].view(np.uint64))
    return int(wsel.sum())


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 1000, 4097, 100003])
def test_selection_projection_int(lib, n):
    rng = np.random.default_rng(n + 5)
    a = gen_col(rng, INT64, n, lo=-30, hi=30)
    b = gen_col(rng, INT64, n, lo=-30, hi=30)
    c = gen_col(rng, UINT64, n, lo=0, hi=50)
    A, B, Cc = Col(0), Col(1), Col(2, "uint")
    filters = [Func("lt", A, B), Func("or", Func("gt", Cc, Const(10, "uint")), Func("isnull", A)),
               Func("in", A, Const(3), Const(None), B, Func("minus", B, Const(5)))]
    projections = [Func("plus", A, B), Func("mul", Cc, Const(3, "uint")), Func("if", Func("ge", A, Const(0)), A, Func("neg", A)),
                   Func("ifnull", B, Const(-1))]
    check([a, b, c], filters[:2], projections)
    check([a, b, c], filters, projections[:2])
    check([a, b, c], [], projections)
    check([a, b, c], filters[:1], [])


@pytest.mark.parametrize("n", [65, 4097, 50000])
def test_selection_projection_real(lib, n):
    rng = np.random.default_rng(n)
    x = gen_col(rng, FLOAT64, n)
    y = gen_col(rng, FLOAT64, n)
    y.values[::7] = 0.0
    x.values[::11] = 0.3       # toBool(real): |f| < 0.5 is "zero"
    k = gen_col(rng, INT64, n, lo=-5, hi=5)
    X, Y, K = Col(0, "real"), Col(1, "real"), Col(2)
    filters = [X, Func("ne", K, Const(0))]
    projections = [Func("div", X, Y), Func("plus", Func("mul", X, Y), Const(1.5, "real")), Func("not", Y), Func("neg", X)]
    nsel = check([x, y, k], filters, projections)
    assert nsel is not None and 0 < nsel < n
    check([x, y, k], [Func("in", X, Const(0.3, "real"), Y)], [Func("minus", X, Y)])


def test_errors_follow_the_narrowed_row_set(lib):
    """an overflowing row raises only while it is still in VecEvalBool's sel slice / in the Selection's output"""
    n = 1000
    big = np.int64(2**62)
    a = Column(INT64, np.full(n, 1, dtype=np.int64), np.ones(n, dtype=bool))
    b = Column(INT64, np.full(n, 2, dtype=np.int64), np.ones(n, dtype=bool))
    a.values[500] = big
    b.values[500] = big            # a + b overflows on row 500 only
    flag = Column(INT64, np.ones(n, dtype=np.int64), np.ones(n, dtype=bool))
    A, B, F = Col(0), Col(1), Col(2)
    ovf = Func("plus", A, B)
    # row 500 is in the set: error, as in the reference
    with pytest.raises(L.TQError) as ei:
        ExprProgram(3, [F], [ovf]).run([a, b, flag])
    assert ei.value.status == L.TQ_ERR_OVERFLOW_BIGINT
    assert check([a, b, flag], [F], [ovf]) is None
    # row 500 filtered out first: the projection never sees it
    flag.values[500] = 0
    assert check([a, b, flag], [F], [ovf]) == n - 1
    # the second CNF item is not evaluated for rows the first one dropped either
    assert check([a, b, flag], [F, Func("gt", ovf, Const(0))], [A]) == n - 1
    # ...but an ETInt NULL in the first item keeps the row in the set (nulls[] quirk): the error comes back
    nn = np.ones(n, dtype=bool)
    nn[500] = False
    flag3 = Column(INT64, np.ones(n, dtype=np.int64), nn)
    assert check([a, b, flag3], [Col(2), Func("gt", ovf, Const(0))], [A]) is None
    # division-by-zero warnings are counted for in-set rows only
    x = Column(FLOAT64, np.ones(n), np.ones(n, dtype=bool))
    z = Column(FLOAT64, np.zeros(n), np.ones(n, dtype=bool))
    keep = Column(INT64, (np.arange(n) % 4 == 0).astype(np.int64), np.ones(n, dtype=bool))
    prog = ExprProgram(3, [Col(2)], [Func("div", Col(0, "real"), Col(1, "real"))])
    outs, sel, warn = prog.run([x, z, keep])
    assert warn == n // 4 and int(sel.sum()) == n // 4
    assert not outs[0].not_null()[sel.astype(bool)].any()   # x / 0 is NULL
    check([x, z, keep], [Col(2)], [Func("div", Col(0, "real"), Col(1, "real"))])


def test_program_is_one_launch_per_slab(lib):
    n = 10000
    rng = np.random.default_rng(1)
    a, b = gen_col(rng, INT64, n, lo=-100, hi=100), gen_col(rng, INT64, n, lo=-100, hi=100)
    prog = ExprProgram(2, [Func("lt", Col(0), Col(1))], [Func("plus", Col(0), Col(1)), Func("minus", Col(0), Col(1))])
    before = lib.tq_kernel_launch_count()
    prog.run([a, b])
    assert lib.tq_kernel_launch_count() - before == 1
