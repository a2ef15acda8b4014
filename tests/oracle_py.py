"""ctypes binding of oracle/liboracle.so (the CPU restatement; TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

from tinysql_b200._lib import TQAggFunc, TQColumn
from tinysql_b200.chunk import Chunk, Column, tq_array, unpack_not_null, _NP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_hash_row.restype = C.c_uint64
        _lib.orc_rowmap_new.restype = C.c_void_p
        _lib.orc_rowmap_get.restype = C.c_int64
        _lib.orc_rowmap_len.restype = C.c_int64
        _lib.orc_mt_join_bench.restype = C.c_int64
        _lib.orc_mt_agg_bench.restype = C.c_int64
        _lib.orc_mt_lt_plus_bench.restype = C.c_int64
    return _lib


def _i32(vals):
    return (C.c_int32 * max(len(vals), 1))(*vals)


def _take(tq_cols, types, n):
    """Copy oracle-malloc'ed columns into numpy-backed Columns."""
    cols = []
    for t, tp in zip(tq_cols, types):
        if n:
            bm = np.ctypeslib.as_array(C.cast(t.null_bitmap, C.POINTER(C.c_uint8)), shape=((n + 7) >> 3,)).copy()
            nn = unpack_not_null(bm, n)
            if tp == 5:  # var-len: offsets + bytes
                off = np.ctypeslib.as_array(C.cast(t.offsets, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
                data = np.ctypeslib.as_array(C.cast(t.data, C.POINTER(C.c_uint8)), shape=(max(int(off[n]), 1),)).copy()
                cols.append(Column(tp, [data[off[i]:off[i + 1]].tobytes() if nn[i] else None for i in range(n)], nn))
            elif tp == 4:  # FLOAT: 4-byte slots
                vals = np.ctypeslib.as_array(C.cast(t.data, C.POINTER(C.c_uint32)), shape=(n,)).copy().view(np.float32)
                cols.append(Column(tp, vals, nn))
            else:
                vals = np.ctypeslib.as_array(C.cast(t.data, C.POINTER(C.c_uint64)), shape=(n,)).copy().view(_NP[tp])
                cols.append(Column(tp, vals, nn))
        else:
            cols.append(Column(tp, [] if tp == 5 else np.zeros(0, dtype=_NP[tp]), np.zeros(0, dtype=bool)))
    return cols


class OrcJoinCond(C.Structure):
    _fields_ = [("op", C.c_int32), ("lhs_col", C.c_int32), ("rhs_col", C.c_int32), ("const_type", C.c_int32), ("const_bits", C.c_uint64)]


def _conds(conds):
    """conds: list of (op, lhs_col, rhs_col) or (op, lhs_col, None, const_type, const_value) over the output row lhs ++ rhs"""
    arr = (OrcJoinCond * max(len(conds), 1))()
    for i, c in enumerate(conds):
        if c[2] is None:
            bits = int(np.array([c[4]], dtype=_NP[c[3]]).view(np.uint64)[0])
            arr[i] = OrcJoinCond(c[0], c[1], -1, c[3], bits)
        else:
            arr[i] = OrcJoinCond(c[0], c[1], c[2], 0, 0)
    return arr


def hash_join(join_type, outer_is_right, build_types, build_cols, probe_types, probe_cols, build_keys, probe_keys, selected=None, conds=(), default_inner=None):
    """default_inner: list of per-inner-column values (None = NULL) — PhysicalHashJoin.DefaultValues"""
    lib = load()
    ncols = len(build_cols) + len(probe_cols)
    out = (TQColumn * ncols)()
    n = C.c_int64(0)
    sel = None
    if selected is not None:
        sel = np.ascontiguousarray(selected, dtype=np.uint8)
    dbits = dnn = None
    if default_inner is not None:
        dbits = (C.c_uint64 * len(build_cols))(*[0 if v is None else int(np.array([v], dtype=_NP[t]).view(np.uint64)[0]) for v, t in zip(default_inner, build_types)])
        dnn = (C.c_uint8 * len(build_cols))(*[0 if v is None else 1 for v in default_inner])
    rc = lib.orc_hash_join_full(C.c_int(join_type), C.c_int(1 if outer_is_right else 0), C.c_int(len(build_cols)), _i32(build_types),
                                tq_array(build_cols), C.c_int(len(probe_cols)), _i32(probe_types), tq_array(probe_cols), C.c_int(len(build_keys)),
                                _i32(build_keys), _i32(probe_keys), C.c_void_p(sel.ctypes.data) if sel is not None else None,
                                C.c_int(len(conds)), _conds(conds), dbits, dnn, out, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"oracle join failed: {rc}")
    types = (list(build_types) + list(probe_types)) if outer_is_right else (list(probe_types) + list(build_types))
    cols = _take(out, types, n.value)
    lib.orc_free_columns(C.c_int(ncols), out)
    return Chunk(cols)


def agg_out_type(func, arg_tp):
    if func == 0:
        return 1
    if func in (1, 2) and arg_tp == 2:
        return 1
    if func in (1, 2) and arg_tp == 4:  # SUM / AVG of a FLOAT column: EvalReal -> DOUBLE
        return 3
    return arg_tp


def hash_agg(types, cols, group_by, funcs, n_partial_workers=1):
    """returns (status, Chunk)"""
    lib = load()
    n_rows = cols[0].length if cols else 0
    fa = (TQAggFunc * max(len(funcs), 1))(*[TQAggFunc(f, a) for f, a in funcs])
    out = (TQColumn * max(len(funcs), 1))()
    n = C.c_int64(0)
    rc = lib.orc_hash_agg(C.c_int(len(cols)), _i32(types), tq_array(cols), C.c_int64(n_rows), C.c_int(len(group_by)), _i32(group_by),
                          C.c_int(len(funcs)), fa, C.c_int(n_partial_workers), out, C.byref(n))
    out_types = [agg_out_type(f, types[a] if a >= 0 else 1) for f, a in funcs]
    res = Chunk(_take(out, out_types, n.value if rc == 0 else 0))
    lib.orc_free_columns(C.c_int(len(funcs)), out)
    return rc, res


class OrcAggFinalFunc(C.Structure):
    _fields_ = [("func", C.c_int32), ("arg_col", C.c_int32), ("arg_col2", C.c_int32)]


def cop_partial_agg(types, cols, group_by, funcs):
    """the coprocessor's partial aggregation (mocktikv/aggregate.go): returns (status, out_types, Chunk of partial rows)"""
    lib = load()
    n_rows = cols[0].length if cols else 0
    n_outc = len(group_by) + len(funcs) + sum(1 for f, _ in funcs if f == 2)
    fa = (TQAggFunc * max(len(funcs), 1))(*[TQAggFunc(f, a) for f, a in funcs])
    out = (TQColumn * max(n_outc, 1))()
    ot = (C.c_int * max(n_outc, 1))()
    n = C.c_int64(0)
    rc = lib.orc_cop_partial_agg(C.c_int(len(cols)), _i32(types), tq_array(cols), C.c_int64(n_rows), C.c_int(len(group_by)), _i32(group_by),
                                 C.c_int(len(funcs)), fa, out, ot, C.byref(n))
    out_types = [int(ot[i]) for i in range(n_outc)]
    res = Chunk(_take(out, out_types, n.value if rc == 0 else 0))
    lib.orc_free_columns(C.c_int(n_outc), out)
    return rc, out_types, res


def hash_agg_final(types, cols, group_by, funcs):
    """FinalMode HashAggExec over partial rows; funcs = [(func, arg_col, arg_col2)].  returns (status, Chunk)"""
    lib = load()
    n_rows = cols[0].length if cols else 0
    fa = (OrcAggFinalFunc * max(len(funcs), 1))(*[OrcAggFinalFunc(f, a, b) for f, a, b in funcs])
    out = (TQColumn * max(len(funcs), 1))()
    n = C.c_int64(0)
    rc = lib.orc_hash_agg_final(C.c_int(len(cols)), _i32(types), tq_array(cols), C.c_int64(n_rows), C.c_int(len(group_by)), _i32(group_by),
                                C.c_int(len(funcs)), fa, out, C.byref(n))
    out_types = [1 if f == 0 else types[b if f == 2 else a] for f, a, b in funcs]
    res = Chunk(_take(out, out_types, n.value if rc == 0 else 0))
    lib.orc_free_columns(C.c_int(len(funcs)), out)
    return rc, res


def sort(types, cols, by, limit_offset=0, limit_count=-1):
    """SortExec / TopNExec; by = [(col, desc)].  Ties keep child order."""
    lib = load()
    n_rows = cols[0].length if cols else 0
    out = (TQColumn * max(len(cols), 1))()
    n = C.c_int64(0)
    rc = lib.orc_sort(C.c_int(len(cols)), _i32(types), tq_array(cols), C.c_int64(n_rows), C.c_int(len(by)), _i32([c for c, _ in by]),
                      _i32([1 if d else 0 for _, d in by]), C.c_int64(limit_offset), C.c_int64(limit_count), out, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"oracle sort failed: {rc}")
    res = Chunk(_take(out, types, n.value))
    lib.orc_free_columns(C.c_int(len(cols)), out)
    return res


def merge_join(join_type, outer_is_right, inner_types, inner_cols, outer_types, outer_cols, inner_keys, outer_keys, selected=None, default_inner=None, conds=()):
    """MergeJoinExec over inputs sorted ascending by their keys; output = left ++ right, outer order x inner order"""
    lib = load()
    ncols = len(inner_cols) + len(outer_cols)
    out = (TQColumn * ncols)()
    n = C.c_int64(0)
    sel = np.ascontiguousarray(selected, dtype=np.uint8) if selected is not None else None
    dbits = dnn = None
    if default_inner is not None:
        dbits = (C.c_uint64 * len(inner_cols))(*[0 if v is None else int(np.array([v], dtype=_NP[t]).view(np.uint64)[0]) for v, t in zip(default_inner, inner_types)])
        dnn = (C.c_uint8 * len(inner_cols))(*[0 if v is None else 1 for v in default_inner])
    rc = lib.orc_merge_join(C.c_int(join_type), C.c_int(1 if outer_is_right else 0), C.c_int(len(inner_cols)), _i32(inner_types), tq_array(inner_cols),
                            C.c_int(len(outer_cols)), _i32(outer_types), tq_array(outer_cols), C.c_int(len(inner_keys)), _i32(inner_keys), _i32(outer_keys),
                            C.c_void_p(sel.ctypes.data) if sel is not None else None, C.c_int(len(conds)), _conds(conds), dbits, dnn, out, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"oracle merge join failed: {rc}")
    types = (list(inner_types) + list(outer_types)) if outer_is_right else (list(outer_types) + list(inner_types))
    res = Chunk(_take(out, types, n.value))
    lib.orc_free_columns(C.c_int(ncols), out)
    return res


def str_to_int(b):
    """types.StrToInt in a SELECT statement: (value, ParseInt failed -> ErrOverflow)"""
    v, e = C.c_int64(0), C.c_int(0)
    load().orc_str_to_int(b, C.c_int64(len(b)), C.byref(v), C.byref(e))
    return v.value, bool(e.value)


def vec_filter_string(a):
    """toBool for ETString: (selected bytes, the error of the last non-NULL row)"""
    sel = np.zeros(max(a.length, 1), dtype=np.uint8)
    ta = a.tq()
    e = C.c_int(0)
    load().orc_vec_filter_string(C.c_int64(a.length), C.byref(ta), C.c_void_p(sel.ctypes.data), C.byref(e))
    return sel[: a.length], bool(e.value)


def _vec(fn, out_tp, n, *args):
    out = Column.empty(out_tp, n)
    to = out.tq()
    rc = fn(*args, C.byref(to))
    return rc, out


def vec_compare_int(op, a, b):
    ta, tb = a.tq(), b.tq()
    return _vec(load().orc_vec_compare_int, 1, a.length, C.c_int(op), C.c_int64(a.length), C.byref(ta), C.c_int(a.tp == 2), C.byref(tb), C.c_int(b.tp == 2))


def vec_compare_real(op, a, b):
    ta, tb = a.tq(), b.tq()
    return _vec(load().orc_vec_compare_real, 1, a.length, C.c_int(op), C.c_int64(a.length), C.byref(ta), C.byref(tb))


def vec_arith_int(op, a, b):
    ta, tb = a.tq(), b.tq()
    out_tp = 2 if (a.tp == 2 or b.tp == 2) else 1
    return _vec(load().orc_vec_arith_int, out_tp, a.length, C.c_int(op), C.c_int64(a.length), C.byref(ta), C.c_int(a.tp == 2), C.byref(tb), C.c_int(b.tp == 2))


def vec_arith_real(op, a, b):
    ta, tb = a.tq(), b.tq()
    out = Column.empty(3, a.length)
    to = out.tq()
    dz = C.c_int64(0)
    rc = load().orc_vec_arith_real(C.c_int(op), C.c_int64(a.length), C.byref(ta), C.byref(tb), C.byref(to), C.byref(dz))
    return rc, out, dz.value


def vec_logic(op, a, b):
    ta, tb = a.tq(), b.tq()
    return _vec(load().orc_vec_logic, 1, a.length, C.c_int(op), C.c_int64(a.length), C.byref(ta), C.byref(tb))


def vec_unary(op, a):
    ta = a.tq()
    return _vec(load().orc_vec_unary, 3 if op == 3 else 1, a.length, C.c_int(op), C.c_int64(a.length), C.byref(ta), C.c_int(a.tp == 2))


def vec_if(c, a, b):
    tc, ta, tb = c.tq(), a.tq(), b.tq()
    return _vec(load().orc_vec_if, a.tp, a.length, C.c_int64(a.length), C.byref(tc), C.byref(ta), C.byref(tb))


def vec_ifnull(a, b):
    ta, tb = a.tq(), b.tq()
    return _vec(load().orc_vec_ifnull, a.tp, a.length, C.c_int64(a.length), C.byref(ta), C.byref(tb))


def vec_in_int(a, lst):
    ta = a.tq()
    arr = tq_array(lst)
    return _vec(load().orc_vec_in_int, 1, a.length, C.c_int64(a.length), C.byref(ta), C.c_int(a.tp == 2), C.c_int(len(lst)), arr,
                _i32([1 if c.tp == 2 else 0 for c in lst]))


def vec_compare_string(op, a, b):
    ta, tb = a.tq(), b.tq()
    return _vec(load().orc_vec_compare_string, 1, a.length, C.c_int(op), C.c_int64(a.length), C.byref(ta), C.byref(tb))


def vec_string_unary(op, a):
    ta = a.tq()
    return _vec(load().orc_vec_string_unary, 1, a.length, C.c_int(op), C.c_int64(a.length), C.byref(ta))


def vec_filter_int(a):
    ta = a.tq()
    sel = np.zeros(max(a.length, 1), dtype=np.uint8)
    load().orc_vec_filter_int(C.c_int64(a.length), C.byref(ta), C.c_void_p(sel.ctypes.data))
    return sel[: a.length]


def vec_in_real(a, lst):
    ta = a.tq()
    return _vec(load().orc_vec_in_real, 1, a.length, C.c_int64(a.length), C.byref(ta), C.c_int(len(lst)), tq_array(lst))


def vec_in_string(a, lst):
    ta = a.tq()
    return _vec(load().orc_vec_in_string, 1, a.length, C.c_int64(a.length), C.byref(ta), C.c_int(len(lst)), tq_array(lst))


def vec_pick_string(mode, cond, a, b):
    """mode 0: IF(cond, a, b); mode 1: IFNULL(a, b) over var-len columns"""
    out = (TQColumn * 1)()
    tc = cond.tq() if cond is not None else TQColumn()
    ta, tb = a.tq(), b.tq()
    rc = load().orc_vec_pick_string(C.c_int(mode), C.c_int64(a.length), C.byref(tc), C.byref(ta), C.byref(tb), out)
    res = _take(out, [5], a.length)[0]
    load().orc_free_columns(C.c_int(1), out)
    return rc, res


def vec_filter_real(a):
    ta = a.tq()
    sel = np.zeros(max(a.length, 1), dtype=np.uint8)
    load().orc_vec_filter_real(C.c_int64(a.length), C.byref(ta), C.c_void_p(sel.ctypes.data))
    return sel[: a.length]
