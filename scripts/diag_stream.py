"""diagnostic: which rows does the streaming join lose / duplicate?"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import INT64, Column, DeviceColumn, device_to_host
lib = L.load(); L.check(lib.tq_init(0))
nb, npr = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(3)
bk = rng.permutation(nb).astype(np.int64)
pk = rng.integers(0, nb, npr).astype(np.int64)
d_b = [DeviceColumn.from_host(Column(INT64, x)) for x in (bk, bk * 7 + 1)]
d_p = [DeviceColumn.from_host(Column(INT64, x)) for x in (pk, np.arange(npr))]
t = (C.c_int32 * 2)(1, 1); k = (C.c_int32 * 1)(0)
d = L.TQJoinDesc(0, 1, 2, t, 2, t, 1, k, k, 0, 0)
h = C.c_void_p(); L.check(lib.tq_join_create(C.byref(d), C.byref(h)))
def arr(cols, n):
    a = (L.TQColumn * 2)()
    for i, c in enumerate(cols):
        a[i].length, a[i].data, a[i].null_bitmap, a[i].offsets = n, c._data.value, None, None
    return a
L.check(lib.tq_join_put_build(h, arr(d_b, nb), 1)); L.check(lib.tq_join_finalize_build(h))
L.check(lib.tq_join_put_probe(h, arr(d_p, npr), None, 1)); L.check(lib.tq_join_probe_eof(h))
out = (L.TQColumn * 4)(); n, eof = C.c_int64(0), C.c_int32(0)
L.check(lib.tq_join_next_device(h, out, C.byref(n), C.byref(eof)))
cols = [device_to_host(INT64, out[c].data, None, n.value).values for c in range(4)]
st = (C.c_int64 * 8)(); lib.tq_join_stats(h, st)
print("rows", n.value, "parts", st[2], "env", {k: v for k, v in os.environ.items() if k.startswith("TQ_")})
ids = cols[3]
cnt = np.bincount(ids[(ids >= 0) & (ids < npr)], minlength=npr)
dup = np.nonzero(cnt > 1)[0]; miss = np.nonzero(cnt == 0)[0]
print("dup ids", dup.size, "missing ids", miss.size, "out-of-range", int(((ids < 0) | (ids >= npr)).sum()))
print("B.k==P.k", bool(np.array_equal(cols[0], cols[2])), "B.v ok", bool(np.array_equal(cols[1], cols[0] * 7 + 1)), "P.k==pk[id]", float((cols[2] == pk[np.clip(ids, 0, npr - 1)]).mean()))
if dup.size:
    pos_dup = np.nonzero(np.isin(ids, dup[:2000]))[0]
    print("positions of some dup rows: min", pos_dup.min(), "max", pos_dup.max(), "first", pos_dup[:12])
    print("missing sample", miss[:12], "dup sample", dup[:12])
    # missing rows: tile index in the probe input (4096 / 2048 rows) and position inside it
    print("missing row %4096 hist (first 16 bins of 256)", np.bincount((miss % 4096) // 256, minlength=16))
    print("missing tile ids", np.unique(miss // 4096)[:20], "n tiles with misses", np.unique(miss // 4096).size)
L.check(lib.tq_join_destroy(h))
