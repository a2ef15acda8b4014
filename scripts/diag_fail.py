"""Diagnostics for the two failing GPU parity tests: prints the rows that differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import FLOAT64, INT64, Column
from tinysql_b200.executor import INNER_JOIN, LEFT_OUTER_JOIN, RIGHT_OUTER_JOIN, HashJoinExec, MockDataSource
from util import canon, gen_col
L.check(L.load().tq_init(0))


def diff(got, want, show=6):
    cg, cw = canon(got), canon(want)
    print("  rows got", cg.shape, "want", cw.shape)
    from collections import Counter
    a = Counter(map(bytes, cg)); b = Counter(map(bytes, cw))
    only_g = list((a - b).elements()); only_w = list((b - a).elements())
    print("  only in got:", len(only_g), " only in want:", len(only_w))
    w = cg.shape[1]
    for name, rows in (("got", only_g), ("want", only_w)):
        for r in rows[:show]:
            print("   ", name, np.frombuffer(r, dtype=np.uint64).reshape(-1, 2).tolist() if w % 2 == 0 else np.frombuffer(r, dtype=np.uint64).tolist())
    return only_g, only_w


which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("both", "default"):
    for jt, oir in ((LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)):
        nb, npr = 300, 2000
        rng = np.random.default_rng(nb + jt)
        bcols = [gen_col(rng, INT64, nb, 0.05, 0, nb // 2 + 2), gen_col(rng, INT64, nb, 0.1, -50, 50), gen_col(rng, FLOAT64, nb, 0.1)]
        pcols = [gen_col(rng, INT64, npr, 0.05, 0, nb), gen_col(rng, INT64, npr, 0.1, -50, 50)]
        bt, pt = [INT64, INT64, FLOAT64], [INT64, INT64]
        defaults = [None, 0, 2.5]
        for conds in ((), ([(0, 1, 3)] if oir else [(0, 3, 1)])):
            e = HashJoinExec(MockDataSource(pt, pcols, 1 << 16), MockDataSource(bt, bcols, 1 << 16), [0], [0], jt, oir, None, 1 << 18, other_conditions=conds, default_inner=defaults)
            e.Open(); got = e.drain(); e.Close()
            want = O.hash_join(jt, oir, bt, bcols, pt, pcols, [0], [0], None, conds, default_inner=defaults)
            print("default_inner jt", jt, "oir", oir, "conds", conds)
            diff(got, want)
if which in ("both", "shapes"):
    nbc, npc = 4, 4
    rng = np.random.default_rng(nbc * 10 + npc)
    nb, npr = 400000, 1500000
    s = np.int64(np.uint64(0xA5C3F00DDEADBEEF).astype(np.int64))
    bk = rng.permutation(nb * 2)[:nb].astype(np.int64)
    bk[7] = s
    bcols = [Column(INT64, bk)] + [Column(INT64, bk * (c + 3) + c) for c in range(1, nbc)]
    pk = rng.integers(0, nb * 2, npr).astype(np.int64)
    pk[11] = s
    pcols = [Column(FLOAT64, rng.random(npr)) for _ in range(npc - 1)] + [Column(INT64, pk)]
    for variant in ("with-marker", "no-marker"):
        if variant == "no-marker":
            bk2 = bk.copy(); bk2[7] = 5_000_000
            bcols = [Column(INT64, bk2)] + [Column(INT64, bk2 * (c + 3) + c) for c in range(1, nbc)]
        e = HashJoinExec(MockDataSource([FLOAT64] * 3 + [INT64], pcols, 1 << 20), MockDataSource([INT64] * 4, bcols, 1 << 20), [3], [0], INNER_JOIN, True, None, 0)
        e.Open(); got = e.drain(); e.Close()
        want = O.hash_join(INNER_JOIN, True, [INT64] * 4, bcols, [FLOAT64] * 3 + [INT64], pcols, [0], [3], None)
        print("shapes[4-4]", variant, os.environ.get("TQ_JOIN_OLD_FAST"), os.environ.get("TQ_JOIN_PP_VARIANT"))
        og, ow = diff(got, want, 4)
        if ow:
            keys = np.array([np.frombuffer(r, dtype=np.uint64)[1] for r in ow]).astype(np.int64)
            print("  missing build keys (first)", keys[:10], " probe positions of these keys:", [int(np.nonzero(pk == k)[0][0]) for k in keys[:10]])
