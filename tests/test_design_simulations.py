"""CPU suite: numpy simulations of device algorithms whose CUDA implementations have not run on a GPU yet (round-1
experiments).  They pin the ALGORITHM — the same steps, the same intermediate arrays — against the oracle, so that a
round-2 failure can only be a CUDA-level slip, not a design error.
  * OtherConditions post-filter (tinysql_b200/csrc/othercond.cu: k_oc_eval / k_oc_decide / k_oc_compact)
  * run placement of the TMA bulk-store push kernel (join.cu k_push_bulk: parity padding, head / mid / tail split)"""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200.chunk import INT64, Column
from util import gen_col

INNER, LEFT = 0, 1


def oc_filter_sim(rows_data, rows_nn, n_probe, outer, build_key_col, rowid_col, build_lo, build_hi, conds):
    """rows_data / rows_nn: [ncols][n] arrays of the unfiltered join result (incl. the hidden row-id column for outer joins)"""
    n = rows_data[0].shape[0]
    flag = np.zeros(n, dtype=np.uint8)
    surv = np.zeros(n_probe, dtype=np.int64)
    first = np.full(n_probe, np.iinfo(np.int64).max)
    for i in range(n):                                            # k_oc_eval
        match = rows_nn[build_key_col][i] if outer else True
        if not match:
            continue
        ok = True
        for op, a, b, k in conds:
            if not rows_nn[a][i] or (b is not None and not rows_nn[b][i]):
                ok = False
                break
            x, y = int(rows_data[a][i]), (int(rows_data[b][i]) if b is not None else k)
            if not [x < y, x <= y, x > y, x >= y, x == y, x != y][op]:
                ok = False
                break
        flag[i] = 1 if ok else 2
        if outer:
            pid = int(rows_data[rowid_col][i])
            if ok:
                surv[pid] += 1
            else:
                first[pid] = min(first[pid], i)
    keep = np.zeros(n, dtype=bool)
    for i in range(n):                                            # k_oc_decide
        keep[i] = flag[i] in (0, 1)
        if flag[i] == 2 and outer:
            pid = int(rows_data[rowid_col][i])
            if surv[pid] == 0 and first[pid] == i:
                flag[i] = 3
                keep[i] = True
    out = []
    for i in range(n):                                            # k_oc_compact
        if not keep[i]:
            continue
        row = []
        for c in range(len(rows_data)):
            nn = rows_nn[c][i] and not (flag[i] == 3 and build_lo <= c < build_hi)
            row.append(int(rows_data[c][i]) if nn else None)
        out.append(tuple(row))
    return out


@pytest.mark.parametrize("jt", [INNER, LEFT])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_other_conditions_post_filter_algorithm(jt, seed):
    rng = np.random.default_rng(seed)
    nb, npr = 60, 300
    b = [gen_col(rng, INT64, nb, 0.05, 0, 20), gen_col(rng, INT64, nb, 0.1, -5, 5)]
    p = [gen_col(rng, INT64, npr, 0.05, 0, 24), gen_col(rng, INT64, npr, 0.1, -5, 5)]
    rowid = Column(INT64, np.arange(npr))
    # probe side is the left child: result = p0 p1 [rowid] b0 b1
    conds_user = [(0, 3, 1, 0), (5, 1, None, 2)]                  # b1 < p1 and p1 != 2   (user columns p0 p1 b0 b1)
    want = O.hash_join(jt, False, [INT64, INT64], b, [INT64, INT64], p, [0], [0], None,
                       [(op, a, bb) if bb is not None else (op, a, None, INT64, k) for op, a, bb, k in conds_user]).rows()
    # the unfiltered device result, with the hidden row-id column the device path adds on the probe side
    raw = O.hash_join(jt, False, [INT64, INT64], b, [INT64, INT64, INT64], p + [rowid], [0], [0])
    data = [c.values for c in raw.cols]
    nn = [c.not_null() for c in raw.cols]
    shift = lambda u: u if u < 2 else u + 1                        # user column -> column of the batch with the hidden one
    conds = [(op, shift(a), None if bb is None else shift(bb), k) for op, a, bb, k in conds_user]
    got = oc_filter_sim(data, nn, npr, jt != INNER, build_key_col=3, rowid_col=2, build_lo=3, build_hi=5, conds=conds)
    got = [(r[0], r[1], r[3], r[4]) for r in got]                  # drop the hidden column
    key = lambda r: tuple((0, 0) if v is None else (1, v) for v in r)
    assert sorted(got, key=key) == sorted(want, key=key)


@pytest.mark.parametrize("seed", range(20))
def test_push_bulk_run_placement(seed):
    """k_push_bulk: every run sits in the stage with the parity of its destination row; head / mid / tail cover the run
    exactly once; the bulk part is 16-byte aligned on both sides; runs never overlap; the stage never overflows"""
    rng = np.random.default_rng(seed)
    n_parts = int(rng.integers(1, 9))
    tile = 4096
    pid = rng.integers(0, n_parts, tile)
    cnt = np.bincount(pid, minlength=n_parts)
    g = rng.integers(0, 1 << 20, n_parts)                          # claimed first destination rows
    cur, start = 0, []
    for bn in range(n_parts):
        st = cur + ((cur ^ int(g[bn])) & 1)
        start.append(st)
        cur = st + int(cnt[bn])
    assert cur <= tile + 16
    used = np.zeros(tile + 16, dtype=int)
    for bn in range(n_parts):
        st, c, gg = start[bn], int(cnt[bn]), int(g[bn])
        assert c == 0 or (st & 1) == (gg & 1)
        head = 1 if (c and st & 1) else 0
        mid = (c - head) & ~1
        tail = c - head - mid
        assert head + mid + tail == c and tail in (0, 1)
        if mid:
            assert (st + head) % 2 == 0 and (gg + head) % 2 == 0    # 16-byte aligned source and destination (8-byte rows)
        covered = ([st] if head else []) + list(range(st + head, st + head + mid)) + ([st + c - 1] if tail else [])
        assert covered == list(range(st, st + c))
        dest = ([gg] if head else []) + list(range(gg + head, gg + head + mid)) + ([gg + c - 1] if tail else [])
        assert dest == list(range(gg, gg + c))
        used[st: st + c] += 1
    assert used.max() <= 1
