"""GPU parity for device-resident operator chains (csrc/sort.cu: TQ_MEM_DEVICE chunks into tq_sort_put / tq_mjoin_put_*, results
lent by tq_sort_next_device / tq_mjoin_next_device): rows produced on the device — here by the hash join — are sorted / merge-joined
without visiting the host, and the result is what the oracle computes for the same chain.  (The file name keeps these tests last.)"""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import FLOAT32, FLOAT64, INT64, UINT64, Chunk, Column, DeviceColumn, device_to_host, tq_array
from util import assert_same_multiset, assert_same_ordered, gen_col

pytestmark = pytest.mark.gpu


def dev_chunks(cols, piece):
    """yield (device columns, tq_column array) per piece of `piece` rows"""
    n = cols[0].length
    for lo in range(0, max(n, 1), piece):
        hi = min(lo + piece, n)
        dcs = [DeviceColumn.from_host(Column(c.tp, c.values[lo:hi], c.not_null()[lo:hi] if c.bitmap is not None else None)) for c in cols]
        arr = (L.TQColumn * len(dcs))(*[d.tq() for d in dcs])
        for i, c in enumerate(cols):
            if c.bitmap is None:
                arr[i].null_bitmap = None
        yield dcs, arr


def lent_to_host(types, out, n):
    return Chunk([device_to_host(t, out[i].data, out[i].null_bitmap, n) for i, t in enumerate(types)])


def sort_device(lib, types, cols, by, piece, off=0, cnt=-1, host_next=False):
    d = L.TQSortDesc(len(types), (C.c_int32 * len(types))(*types), len(by), (C.c_int32 * max(len(by), 1))(*[c for c, _ in by]),
                     (C.c_int32 * max(len(by), 1))(*[1 if x else 0 for _, x in by]), off, cnt)
    h = C.c_void_p()
    L.check(lib.tq_sort_create(C.byref(d), C.byref(h)))
    keep = []
    for dcs, arr in dev_chunks(cols, piece):
        L.check(lib.tq_sort_put(h, arr, L.TQ_MEM_DEVICE))
        for dc in dcs:
            dc.free()            # the call has copied what it keeps
    L.check(lib.tq_sort_eof(h))
    n, eof = C.c_int64(0), C.c_int32(0)
    if host_next:
        res = [Column.empty(t, 1 << 20) for t in types]
        L.check(lib.tq_sort_next(h, 1 << 20, tq_array(res, 1 << 20), C.byref(n), C.byref(eof)))
        got = Chunk([Column(c.tp, c.values[: n.value].copy(), c.not_null()[: n.value].copy()) for c in res])
    else:
        out = (L.TQColumn * len(types))()
        L.check(lib.tq_sort_next_device(h, out, C.byref(n), C.byref(eof)))
        got = lent_to_host(types, out, n.value)
        L.check(lib.tq_sort_next_device(h, out, C.byref(n), C.byref(eof)))
        assert n.value == 0 and eof.value == 1
    L.check(lib.tq_sort_destroy(h))
    return got


@pytest.mark.parametrize("n,piece", [(0, 1000), (1, 1000), (70001, 9000), (70001, 100000)])
def test_sort_device_chunks(lib, n, piece):
    rng = np.random.default_rng(n + piece)
    types = [INT64, UINT64, FLOAT64, INT64]
    cols = [gen_col(rng, INT64, n, 0.1, -50, 50), gen_col(rng, UINT64, n, 0.0, 0, 7), Column(FLOAT64, rng.integers(-9, 9, n) * 0.5, rng.random(n) > 0.1),
            Column(INT64, np.arange(n))]
    for by in ([(0, False)], [(2, True), (1, False)], []):
        assert_same_ordered(sort_device(lib, types, cols, by, piece), O.sort(types, cols, by))
    assert_same_ordered(sort_device(lib, types, cols, [(0, True), (2, False)], piece, 5, 1000), O.sort(types, cols, [(0, True), (2, False)], 5, 1000))
    assert_same_ordered(sort_device(lib, types, cols, [(0, False)], piece, host_next=True), O.sort(types, cols, [(0, False)]))   # lazy host copy


def arr_nn(dcs):
    a = (L.TQColumn * len(dcs))(*[d.tq() for d in dcs])
    for i in range(len(dcs)):
        a[i].null_bitmap = None
    return a


def chain_after_join(lib, joined, release_join, want_join, types, b):
    """joined: 4 device tq_columns (B.k, B.v, P.k, P.v) lent by a producer; release_join() gives them back once the first
    consumer holds its own copy.  SortExec -> SortExec -> MergeJoinExec, every hand-over in HBM."""
    n, eof, ns, nm = C.c_int64(0), C.c_int32(0), C.c_int64(0), C.c_int64(0)
    jt = [INT64] * 4
    # ---- SortExec over the producer's device rows: ORDER BY B.v DESC, P.v
    sd = L.TQSortDesc(4, (C.c_int32 * 4)(*jt), 2, (C.c_int32 * 2)(1, 3), (C.c_int32 * 2)(1, 0), 0, -1)
    hs = C.c_void_p()
    L.check(lib.tq_sort_create(C.byref(sd), C.byref(hs)))
    L.check(lib.tq_sort_put(hs, joined, L.TQ_MEM_DEVICE))
    release_join()                                                           # the sort has its own copy
    L.check(lib.tq_sort_eof(hs))
    sorted_cols = (L.TQColumn * 4)()
    L.check(lib.tq_sort_next_device(hs, sorted_cols, C.byref(ns), C.byref(eof)))
    got_sorted = lent_to_host(jt, sorted_cols, ns.value)
    assert_same_ordered(got_sorted, O.sort(jt, want_join.cols, [(1, True), (3, False)]))   # P.v is unique: the order is total
    # ---- a second SortExec fed by the first one's lent columns: ORDER BY B.k, P.v
    sk = L.TQSortDesc(4, (C.c_int32 * 4)(*jt), 2, (C.c_int32 * 2)(0, 3), (C.c_int32 * 2)(0, 0), 0, -1)
    hk = C.c_void_p()
    L.check(lib.tq_sort_create(C.byref(sk), C.byref(hk)))
    L.check(lib.tq_sort_put(hk, sorted_cols, L.TQ_MEM_DEVICE))
    L.check(lib.tq_sort_destroy(hs))
    L.check(lib.tq_sort_eof(hk))
    by_key = (L.TQColumn * 4)()
    L.check(lib.tq_sort_next_device(hk, by_key, C.byref(ns), C.byref(eof)))
    # ---- MergeJoinExec: inner child = the build table sorted by key (device chunks), outer child = the rows sorted by B.k
    inner_sorted = O.sort(types, b, [(0, False)])
    d_in = [DeviceColumn.from_host(c) for c in inner_sorted.cols]
    md = L.TQMJoinDesc(1, 0, 2, (C.c_int32 * 2)(1, 1), 4, (C.c_int32 * 4)(*jt), 1, (C.c_int32 * 1)(0), (C.c_int32 * 1)(2), None, None)   # left outer, inner.k = outer.P.k
    hm = C.c_void_p()
    L.check(lib.tq_mjoin_create(C.byref(md), C.byref(hm)))
    L.check(lib.tq_mjoin_put_inner(hm, arr_nn(d_in), L.TQ_MEM_DEVICE))
    L.check(lib.tq_mjoin_put_outer(hm, by_key, None, L.TQ_MEM_DEVICE))
    L.check(lib.tq_sort_destroy(hk))
    L.check(lib.tq_mjoin_finish(hm))
    mj_cols = (L.TQColumn * 6)()
    L.check(lib.tq_mjoin_next_device(hm, mj_cols, C.byref(nm), C.byref(eof)))
    got_mj = lent_to_host([INT64] * 6, mj_cols, nm.value)
    outer_sorted = O.sort(jt, want_join.cols, [(0, False), (3, False)])
    assert_same_ordered(got_mj, O.merge_join(1, False, types, inner_sorted.cols, jt, outer_sorted.cols, [0], [2]))
    # the host copy of a device-chunk handle is made on demand and equals the lent one
    res = [Column.empty(INT64, max(nm.value, 1)) for _ in range(6)]
    L.check(lib.tq_mjoin_next(hm, max(nm.value, 1), tq_array(res, max(nm.value, 1)), C.byref(n), C.byref(eof)))
    assert n.value == nm.value
    assert_same_ordered(Chunk([Column(INT64, c.values[: n.value], c.not_null()[: n.value]) for c in res]), got_mj)
    L.check(lib.tq_mjoin_destroy(hm))
    for dc in d_in:
        dc.free()


def chain_tables(nb, npr):
    rng = np.random.default_rng(12)
    b = [Column(INT64, rng.permutation(nb)), gen_col(rng, INT64, nb, 0.0, 0, 100)]
    p = [gen_col(rng, INT64, npr, 0.0, 0, nb + nb // 10), Column(INT64, np.arange(npr))]
    return [INT64, INT64], b, p


def test_join_then_sort_then_merge_join_stay_on_the_device(lib):
    types, b, p = chain_tables(40000, 300000)
    d_b, d_p = [DeviceColumn.from_host(c) for c in b], [DeviceColumn.from_host(c) for c in p]
    t, k = (C.c_int32 * 2)(1, 1), (C.c_int32 * 1)(0)
    jd = L.TQJoinDesc(0, 1, 2, t, 2, t, 1, k, k, 0, 0)
    hj = C.c_void_p()
    L.check(lib.tq_join_create(C.byref(jd), C.byref(hj)))
    L.check(lib.tq_join_put_build(hj, arr_nn(d_b), L.TQ_MEM_DEVICE))
    L.check(lib.tq_join_finalize_build(hj))
    L.check(lib.tq_join_put_probe(hj, arr_nn(d_p), None, L.TQ_MEM_DEVICE))
    L.check(lib.tq_join_probe_eof(hj))
    joined = (L.TQColumn * 4)()
    n, eof = C.c_int64(0), C.c_int32(0)
    L.check(lib.tq_join_next_device(hj, joined, C.byref(n), C.byref(eof)))
    want_join = O.hash_join(0, True, types, b, types, p, [0], [0])           # (B.k, B.v, P.k, P.v)
    assert n.value == want_join.num_rows()
    chain_after_join(lib, joined, lambda: L.check(lib.tq_join_destroy(hj)), want_join, types, b)
    for dc in d_b + d_p:
        dc.free()


def test_device_chunk_rules(lib):
    types = [INT64, FLOAT32]
    d = L.TQSortDesc(2, (C.c_int32 * 2)(*types), 1, (C.c_int32 * 1)(0), (C.c_int32 * 1)(0), 0, -1)
    h = C.c_void_p()
    L.check(lib.tq_sort_create(C.byref(d), C.byref(h)))
    dc = DeviceColumn.from_host(Column(INT64, [3, 1, 2]))
    a = (L.TQColumn * 2)(dc.tq(), dc.tq())
    assert lib.tq_sort_put(h, a, L.TQ_MEM_DEVICE) == L.TQ_ERR_UNSUPPORTED_TYPE        # FLOAT (4-byte) columns come as host chunks
    L.check(lib.tq_sort_destroy(h))
    d = L.TQSortDesc(1, (C.c_int32 * 1)(INT64), 1, (C.c_int32 * 1)(0), (C.c_int32 * 1)(0), 0, -1)
    L.check(lib.tq_sort_create(C.byref(d), C.byref(h)))
    a1 = (L.TQColumn * 1)(dc.tq())
    L.check(lib.tq_sort_put(h, a1, L.TQ_MEM_DEVICE))
    host = Column(INT64, [9, 8])
    assert lib.tq_sort_put(h, tq_array([host]), L.TQ_MEM_HOST) == L.TQ_ERR_STATE       # not both kinds into one handle
    L.check(lib.tq_sort_eof(h))
    out = (L.TQColumn * 1)()
    n, eof = C.c_int64(0), C.c_int32(0)
    L.check(lib.tq_sort_next_device(h, out, C.byref(n), C.byref(eof)))
    assert list(device_to_host(INT64, out[0].data, out[0].null_bitmap, n.value).values) == [1, 2, 3]
    L.check(lib.tq_sort_destroy(h))
    dc.free()
