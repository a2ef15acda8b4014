"""expression.vecEval* surface (expression/builtin.go:256-263) bound to the CUDA kernels.

Each function mirrors one vectorized builtin signature: it takes evaluated argument columns and
returns the result column, raising TQError with the reference's error kinds (types.ErrOverflow…).
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .chunk import BYTES, Column, FLOAT64, INT64, UINT64, VarColumn, tq_array

LT, LE, GT, GE, EQ, NE = range(6)
PLUS, MINUS, MUL, DIV = range(4)
AND, OR = 0, 1
NOT_INT, NOT_REAL, MINUS_INT, MINUS_REAL, ISNULL = range(5)
STRCMP = 6                     # tq_vec_compare_string op beyond LT..NE
STR_LENGTH, STR_ISNULL = 0, 1  # tq_vec_string_unary ops


def _u(col):
    return 1 if col.tp == UINT64 else 0


def vec_compare_int(op, a, b):
    """builtin{LT,LE,GT,GE,EQ,NE}IntSig.vecEvalInt — expression/builtin_compare_vec.go:22-292"""
    out = Column.empty(INT64, a.length)
    ta, tb, to = a.tq(), b.tq(), out.tq()
    L.check(L.load().tq_vec_compare_int(op, a.length, C.byref(ta), _u(a), C.byref(tb), _u(b), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_compare_real(op, a, b):
    """builtin{LT..NE}RealSig.vecEvalInt — expression/builtin_compare_vec_generated.go"""
    out = Column.empty(INT64, a.length)
    ta, tb, to = a.tq(), b.tq(), out.tq()
    L.check(L.load().tq_vec_compare_real(op, a.length, C.byref(ta), C.byref(tb), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_compare_string(op, a, b):
    """builtin{LT..NE}StringSig.vecEvalInt (builtin_compare_vec_generated.go:65-555); op STRCMP: builtinStrcmpSig
    (builtin_string_vec.go:52-83).  a, b: var-len columns."""
    out = Column.empty(INT64, a.length)
    ta, tb, to = a.tq(), b.tq(), out.tq()
    L.check(L.load().tq_vec_compare_string(op, a.length, C.byref(ta), C.byref(tb), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_string_unary(op, a):
    """STR_LENGTH: builtinLengthSig (builtin_string.go:75-81); STR_ISNULL: builtinStringIsNullSig (builtin_string_vec.go:21-42)"""
    out = Column.empty(INT64, a.length)
    ta, to = a.tq(), out.tq()
    L.check(L.load().tq_vec_string_unary(op, a.length, C.byref(ta), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_arith_int(op, a, b):
    """builtinArithmetic{Plus,Minus,Multiply}IntSig.vecEvalInt — expression/builtin_arithmetic_vec.go"""
    out = Column.empty(UINT64 if (_u(a) or _u(b)) else INT64, a.length)
    ta, tb, to = a.tq(), b.tq(), out.tq()
    L.check(L.load().tq_vec_arith_int(op, a.length, C.byref(ta), _u(a), C.byref(tb), _u(b), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_arith_real(op, a, b):
    """builtinArithmetic{Plus,Minus,Multiply,Divide}RealSig.vecEvalReal; returns (column, div-by-zero warnings)"""
    out = Column.empty(FLOAT64, a.length)
    ta, tb, to = a.tq(), b.tq(), out.tq()
    dz = C.c_int64(0)
    L.check(L.load().tq_vec_arith_real(op, a.length, C.byref(ta), C.byref(tb), C.byref(to), C.byref(dz), L.TQ_MEM_HOST))
    return out, dz.value


def vec_logic(op, a, b):
    """builtinLogic{And,Or}Sig.vecEvalInt — expression/builtin_op_vec.go:29-68,173-215"""
    out = Column.empty(INT64, a.length)
    ta, tb, to = a.tq(), b.tq(), out.tq()
    L.check(L.load().tq_vec_logic(op, a.length, C.byref(ta), C.byref(tb), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_unary(op, a):
    """UnaryNot / UnaryMinus / IsNull — expression/builtin_op_vec.go"""
    out_tp = FLOAT64 if op == MINUS_REAL else INT64
    out = Column.empty(out_tp, a.length)
    ta, to = a.tq(), out.tq()
    L.check(L.load().tq_vec_unary(op, a.length, C.byref(ta), _u(a), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_if(cond, a, b):
    """builtinIf{Int,Real}Sig — expression/builtin_control_vec_generated.go:117-207"""
    out = Column.empty(a.tp, a.length)
    tc, ta, tb, to = cond.tq(), a.tq(), b.tq(), out.tq()
    L.check(L.load().tq_vec_if(a.length, C.byref(tc), C.byref(ta), C.byref(tb), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_ifnull(a, b):
    """builtinIfNull{Int,Real}Sig — expression/builtin_control_vec_generated.go:23-79"""
    out = Column.empty(a.tp, a.length)
    ta, tb, to = a.tq(), b.tq(), out.tq()
    L.check(L.load().tq_vec_ifnull(a.length, C.byref(ta), C.byref(tb), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_in_int(a, lst):
    """builtinInIntSig — expression/builtin_other_vec_generated.go:24-96"""
    out = Column.empty(INT64, a.length)
    ta, to = a.tq(), out.tq()
    arr = tq_array(lst)
    flags = (C.c_int32 * max(len(lst), 1))(*[_u(c) for c in lst])
    L.check(L.load().tq_vec_in_int(a.length, C.byref(ta), _u(a), len(lst), arr, flags, C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_lt_plus_int(a, b):
    """BASELINE config 2 in one pass: (a < b, a + b) over signed BIGINT columns."""
    lt, plus = Column.empty(INT64, a.length), Column.empty(INT64, a.length)
    ta, tb, t1, t2 = a.tq(), b.tq(), lt.tq(), plus.tq()
    L.check(L.load().tq_vec_lt_plus_int(a.length, C.byref(ta), C.byref(tb), C.byref(t1), C.byref(t2), L.TQ_MEM_HOST))
    return lt, plus


def vectorized_filter(a):
    """expression.VectorizedFilter over one evaluated int column (expression/chunk_executor.go:196-245)."""
    sel = np.zeros(max(a.length, 1), dtype=np.uint8)
    ta = a.tq()
    L.check(L.load().tq_vec_filter_int(a.length, C.byref(ta), sel.ctypes.data, L.TQ_MEM_HOST))
    return sel[: a.length]


def vectorized_filter_real(a):
    """VecEvalBool / toBool for an ETReal expression (expression/expression.go:296-307): zero iff RoundFloat(f) == 0."""
    sel = np.zeros(max(a.length, 1), dtype=np.uint8)
    ta = a.tq()
    L.check(L.load().tq_vec_filter_real(a.length, C.byref(ta), sel.ctypes.data, L.TQ_MEM_HOST))
    return sel[: a.length]


def vec_in_real(a, lst):
    """builtinInRealSig — expression/builtin_other_vec_generated.go:151-204"""
    out = Column.empty(INT64, a.length)
    ta, to = a.tq(), out.tq()
    L.check(L.load().tq_vec_in_real(a.length, C.byref(ta), len(lst), tq_array(lst), C.byref(to), L.TQ_MEM_HOST))
    return out


def vec_in_string(a, lst):
    """builtinInStringSig — expression/builtin_other_vec_generated.go:97-149"""
    out = Column.empty(INT64, a.length)
    ta, to = a.tq(), out.tq()
    L.check(L.load().tq_vec_in_string(a.length, C.byref(ta), len(lst), tq_array(lst), C.byref(to), L.TQ_MEM_HOST))
    return out


def _pick_string(fn, n, cap, *args):
    out = VarColumn.empty(BYTES, n, cap)
    to = out.tq()
    L.check(fn(n, *args, C.byref(to), L.TQ_MEM_HOST))
    return out.head(n)


def vec_if_string(cond, a, b):
    """builtinIfStringSig.vecEvalString — expression/builtin_control_vec_generated.go:209-262"""
    tc, ta, tb = cond.tq(), a.tq(), b.tq()
    return _pick_string(L.load().tq_vec_if_string, a.length, int(a.data.size + b.data.size), C.byref(tc), C.byref(ta), C.byref(tb))


def vec_ifnull_string(a, b):
    """builtinIfNullStringSig.vecEvalString — expression/builtin_control_vec_generated.go:81-112"""
    ta, tb = a.tq(), b.tq()
    return _pick_string(L.load().tq_vec_ifnull_string, a.length, int(a.data.size + b.data.size), C.byref(ta), C.byref(tb))
