"""Stress for the (4 build cols, 4 probe cols) streaming join with the empty-marker key on the build side: reports torn rows."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import FLOAT64, INT64, Column
from tinysql_b200.executor import INNER_JOIN, HashJoinExec, MockDataSource
L.check(L.load().tq_init(0))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
marker = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
nbc = npc = 4
rng = np.random.default_rng(44)
nb, npr = 400000, 1500000
s = np.int64(np.uint64(0xA5C3F00DDEADBEEF).astype(np.int64))
bk = rng.permutation(nb * 2)[:nb].astype(np.int64)
if marker:
    bk[7] = s
bcols = [Column(INT64, bk)] + [Column(INT64, bk * (c + 3) + c) for c in range(1, nbc)]
pk = rng.integers(0, nb * 2, npr).astype(np.int64)
pk[11] = s
fl = [rng.random(npr) for _ in range(npc - 1)]
pcols = [Column(FLOAT64, f) for f in fl] + [Column(INT64, pk)]
inb = np.isin(pk, bk)
want_rows = int(inb.sum())
f0_to_row = {v: i for i, v in enumerate(fl[0].view(np.uint64).tolist())}
f2_to_row = {v: i for i, v in enumerate(fl[2].view(np.uint64).tolist())}


def hash_key(k):
    k = np.uint64(k)
    k ^= k >> np.uint64(32)
    k = k * np.uint64(0x9E3779B97F4A7C15)
    return k ^ (k >> np.uint64(32))


for rep in range(reps):
    e = HashJoinExec(MockDataSource([FLOAT64] * 3 + [INT64], pcols, 1 << 20), MockDataSource([INT64] * 4, bcols, 1 << 20), [3], [0], INNER_JOIN, True, None, 0)
    e.Open(); got = e.drain(); e.Close()
    g = [c.values for c in got.cols]          # b0..b3, f0..f2, pk
    n = got.num_rows()
    rows = np.array([f0_to_row.get(v, -1) for v in g[4].view(np.uint64).tolist()])
    unknown = np.nonzero(rows < 0)[0]
    if len(unknown):
        print(f"rep {rep}: {len(unknown)} output rows whose f0 is no probe value; first:", [[hex(int(c.view(np.uint64)[i])) for c in g] for i in unknown[:4]], flush=True)
        rows = np.where(rows < 0, 0, rows)
    bad = np.nonzero((pk[rows] != g[7]) | (g[0] != g[7]) | (fl[1][rows] != g[5]) | (fl[2][rows] != g[6]) | (g[1] != g[0] * 4 + 1))[0]
    uniq = len(np.unique(rows))
    print(f"rep {rep}: rows {n} want {want_rows} distinct probe rows {uniq} torn {len(bad)}", flush=True)
    with np.errstate(over="ignore"):
        for i in bad[:12]:
            a = rows[i]
            b = f2_to_row.get(int(g[6].view(np.uint64)[i]), -1)
            pa, pb = int(hash_key(pk[a]) >> np.uint64(62)), int(hash_key(g[7][i]) >> np.uint64(62))
            print(f"   out slot {i}: floats of probe row {a} (key {pk[a]}, part {pa}, hit {bool(inb[a])}) with key {g[7][i]} of row {b} (part {pb}) delta {b - a}; f1/f2 from A: {fl[1][a] == g[5][i]} {fl[2][a] == g[6][i]} key==pk[B] {b >= 0 and pk[b] == g[7][i]}")
