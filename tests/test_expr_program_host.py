"""Host side of the fused Selection + Projection program (no GPU): lowering and argument validation of tq_expr_eval."""
import ctypes as C

import pytest

from tinysql_b200 import _lib as L
from tinysql_b200.expression import (X_CMP_INT, X_COMPACT, X_CONST, X_FILTER, X_LOGIC, Col, Const, ExprProgram, Func)


def test_lowering_shares_subtrees_and_orders_filters_before_projections():
    a, b = Col(0), Col(1)
    s = Func("plus", a, b)
    prog = ExprProgram(2, [Func("lt", a, b)], [s, Func("mul", s, s)])
    kinds = [o.kind for o in prog.ops]
    assert kinds[:3] == [X_CMP_INT, X_FILTER, X_COMPACT]
    assert len(prog.ops) == 5            # a + b evaluated once
    assert prog.out_regs == [2 + 3, 2 + 4]
    for i, o in enumerate(prog.ops):     # straight-line: operands are inputs or earlier results
        assert max(o.a, o.b, o.c) < 2 + i or o.kind in (X_CONST, X_COMPACT)


def test_in_lowers_to_eq_or_chain():
    prog = ExprProgram(1, [Func("in", Col(0), Const(1), Const(None), Const(7))])
    kinds = [o.kind for o in prog.ops]
    assert kinds.count(X_CMP_INT) == 3 and kinds.count(X_LOGIC) == 2 and kinds[-1] == X_FILTER


def test_limits():
    with pytest.raises(ValueError):
        ExprProgram(9)
    e = Col(0)
    for i in range(40):
        e = Func("plus", e, Const(i))
    with pytest.raises(ValueError):
        ExprProgram(1, [], [e])
    with pytest.raises(ValueError):
        Func("div", Col(0), Col(1))     # integer '/' is decimal division in the reference: not on this path


def _call(n_in, ops, n_out=0, out_regs=None, sel=True):
    lib = L.load()
    arr = (L.TQExprOp * max(len(ops), 1))(*ops)
    cols = (L.TQColumn * 8)()
    outs = (L.TQColumn * 4)()
    regs = (C.c_int32 * 4)(*(out_regs or []))
    buf = (C.c_uint8 * 16)()
    return lib.tq_expr_eval(8, n_in, cols, len(ops), arr, n_out, regs, outs, buf if sel else None, None, L.TQ_MEM_HOST)


def test_abi_rejects_malformed_programs_before_touching_the_device():
    op = L.TQExprOp
    assert _call(9, []) == L.TQ_ERR_INVALID_ARG
    assert _call(2, [op(X_CMP_INT, 0, 0, 2, 0, 0, 0, 0, 0)]) == L.TQ_ERR_INVALID_ARG        # reads its own result
    assert "before it is written" in L.last_error()
    assert _call(2, [op(X_CMP_INT, 9, 0, 1, 0, 0, 0, 0, 0)]) == L.TQ_ERR_INVALID_ARG        # bad operator
    assert _call(2, [op(99, 0, 0, 1, 0, 0, 0, 0, 0)]) == L.TQ_ERR_INVALID_ARG               # bad kind
    assert _call(2, [op(X_CMP_INT, 0, 0, 1, 0, 0, 0, 0, 0)], 1, [7]) == L.TQ_ERR_INVALID_ARG  # output register out of range
    assert _call(2, [], 0, None, sel=False) == L.TQ_ERR_INVALID_ARG                          # neither outputs nor selection
