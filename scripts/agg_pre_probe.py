"""Times the HashAgg update pipeline (tq_agg_stats[2]) for a few group counts; run once with TQ_AGG_NO_PREAGG=1 and once
without to compare the general path with the shared-memory pre-aggregation path."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import FLOAT64, INT64, Column, DeviceColumn

lib = L.load()
L.check(lib.tq_init(0))
n = 50_000_000
x = np.random.default_rng(6).random(n)
dx = DeviceColumn.from_host(Column(FLOAT64, x))
for groups in [int(g) for g in sys.argv[1:]] or [1000, 30000, 1000000]:
    k = np.random.default_rng(5).integers(0, groups, n, dtype=np.int64)
    dk = DeviceColumn.from_host(Column(INT64, k))
    types = (C.c_int32 * 2)(INT64 | 0x100, FLOAT64 | 0x100)
    gb = (C.c_int32 * 1)(0)
    funcs = (L.TQAggFunc * 3)(L.TQAggFunc(1, 1), L.TQAggFunc(0, -1), L.TQAggFunc(5, 0))
    desc = L.TQAggDesc(2, types, 1, gb, 3, funcs, groups)
    best = None
    for it in range(4):
        h = C.c_void_p()
        L.check(lib.tq_agg_create(C.byref(desc), C.byref(h)))
        cols = (L.TQColumn * 2)(dk.tq(), dx.tq())
        cols[0].null_bitmap = None
        cols[1].null_bitmap = None
        L.check(lib.tq_agg_put(h, cols, L.TQ_MEM_DEVICE))
        L.check(lib.tq_agg_eof(h))
        st = (C.c_int64 * 4)()
        lib.tq_agg_stats(h, st)
        L.check(lib.tq_agg_destroy(h))
        if it:
            best = st[2] if best is None else min(best, st[2])
    print(f"preagg={'off' if os.environ.get('TQ_AGG_NO_PREAGG') == '1' else 'on'} groups={groups} rows={n} update_ms={best / 1e6:.3f} launches={st[3]}", flush=True)
    dk.free()
