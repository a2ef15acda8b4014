"""GPU parity for FLOAT (4-byte) and var-len payload columns of HashJoinExec — the reference's own benchmark shape
(executor/benchmark_test.go:391-457: BIGINT key + 5 KiB VARSTRING payload, BASELINE config C1) — against the oracle's
restatement of chunk.CopySelectedJoinRows (util/chunk/chunk_util.go:38-110)."""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, Column
from tinysql_b200.executor import INNER_JOIN, LEFT_OUTER_JOIN, RIGHT_OUTER_JOIN, HashJoinExec, MockDataSource
from util import assert_same_multiset, assert_same_ordered, gen_col

pytestmark = pytest.mark.gpu


def run_join(btypes, bcols, ptypes, pcols, bkeys, pkeys, jt=INNER_JOIN, oir=False, chunk=1024, batch=0, req=None):
    inner, outer = MockDataSource(btypes, bcols, chunk), MockDataSource(ptypes, pcols, chunk)
    e = HashJoinExec(outer, inner, pkeys, bkeys, jt, oir, None, batch, max_chunk_size=req or 1024)
    e.Open()
    got = e.drain()
    e.Close()
    return got, O.hash_join(jt, oir, btypes, bcols, ptypes, pcols, bkeys, pkeys)


def rand_cells(rng, n, max_len, null_frac):
    lens = rng.integers(0, max_len + 1, n)
    nn = rng.random(n) >= null_frac
    return [rng.integers(0, 256, l, dtype=np.uint8).tobytes() if ok else None for l, ok in zip(lens, nn)]


def test_varlen_small_explicit(lib):
    b = [Column(INT64, [1, 2, 3, 2]), Column(BYTES, [b"one", None, b"three" * 1000, b"two-b"]),
         Column(FLOAT32, [1.5, 2.5, 3.5, 4.5], [True, True, False, True])]
    p = [Column(BYTES, [b"p0", b"", b"p2", None]), Column(INT64, [2, 3, 9, 1])]
    for jt, oir in ((INNER_JOIN, False), (INNER_JOIN, True), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)):
        got, want = run_join([INT64, BYTES, FLOAT32], b, [BYTES, INT64], p, [0], [1], jt, oir)
        assert got.rows() == want.rows()


@pytest.mark.parametrize("jt,oir", [(INNER_JOIN, True), (LEFT_OUTER_JOIN, False), (RIGHT_OUTER_JOIN, True)])
def test_varlen_random_batches(lib, jt, oir):
    """both sides carry var-len and FLOAT payloads; ragged chunks, several device batches, duplicate keys, NULL cells"""
    rng = np.random.default_rng(61 + jt)
    nb, npr = 3000, 25000
    bcols = [Column(BYTES, rand_cells(rng, nb, 40, 0.1)), gen_col(rng, INT64, nb, 0.05, 0, 1500),
             Column(FLOAT32, rng.random(nb).astype(np.float32), rng.random(nb) > 0.1)]
    pcols = [gen_col(rng, INT64, npr, 0.05, 0, 2000), Column(FLOAT32, rng.random(npr).astype(np.float32), rng.random(npr) > 0.1),
             Column(BYTES, rand_cells(rng, npr, 24, 0.1)), gen_col(rng, FLOAT64, npr, 0.1)]
    got, want = run_join([BYTES, INT64, FLOAT32], bcols, [INT64, FLOAT32, BYTES, FLOAT64], pcols, [1], [0], jt, oir, chunk=1000, batch=4096)
    assert_same_multiset(got, want)


def test_varlen_c1_benchmark_shape(lib):
    """BASELINE config C1 at 1/5 scale: two tables of (BIGINT k = row, VARSTRING 5 KiB constant), k unique -> every row
    matches once; output = 4 columns.  Checked by properties on the raw result buffers (executor/benchmark_test.go:391-457)."""
    n = 20000
    cell = bytes(range(256)) * 20  # 5 KiB
    k = np.arange(n, dtype=np.int64)
    rng = np.random.default_rng(1)
    build = [Column(INT64, rng.permutation(k)), Column(BYTES, [cell] * n)]
    probe = [Column(INT64, k), Column(BYTES, [cell] * n)]
    inner, outer = MockDataSource([INT64, BYTES], build), MockDataSource([INT64, BYTES], probe)
    e = HashJoinExec(outer, inner, [0], [0], INNER_JOIN, False)
    e.Open()
    rows = 0
    seen = np.zeros(n, dtype=bool)
    while True:
        c = e.Next()
        if c.num_rows() == 0:
            break
        m = c.num_rows()
        rows += m
        assert np.array_equal(c.cols[0].values, c.cols[2].values)
        seen[c.cols[0].values] = True
        for v in (c.cols[1], c.cols[3]):
            assert np.array_equal(v.offsets, np.arange(m + 1) * len(cell))
            assert np.array_equal(v.data.reshape(m, len(cell)), np.broadcast_to(np.frombuffer(cell, dtype=np.uint8), (m, len(cell))))
    e.Close()
    assert rows == n and seen.all()


def test_varlen_partitioned_build(lib):
    """a build side on the partitioned-table path: the payload string of every joined row is a function of its key"""
    rng = np.random.default_rng(9)
    nb, npr = 300000, 700000
    bk = rng.permutation(nb).astype(np.int64)
    cells = [b"k%07d" % v + b"x" * (v % 5) for v in bk]
    build = [Column(INT64, bk), Column(BYTES, cells)]
    pk = rng.integers(0, nb + nb // 4, npr).astype(np.int64)
    probe = [Column(INT64, pk), Column(FLOAT32, (pk % 1000).astype(np.float32))]
    inner, outer = MockDataSource([INT64, BYTES], build, 1 << 18), MockDataSource([INT64, FLOAT32], probe, 1 << 18)
    e = HashJoinExec(outer, inner, [0], [0], LEFT_OUTER_JOIN, False, None, 1 << 18, max_chunk_size=1 << 16)
    e.Open()
    got = e.drain()
    e.Close()
    assert got.num_rows() == npr
    order = np.argsort(got.cols[0].values, kind="stable")
    assert np.array_equal(np.sort(got.cols[0].values), np.sort(pk))
    hit = got.cols[2].not_null()
    assert np.array_equal(hit, got.cols[0].values < nb)
    assert np.array_equal(got.cols[1].values, (got.cols[0].values % 1000).astype(np.float32))
    strs = got.cols[3].tolist()
    keys = got.cols[0].values
    for i in rng.integers(0, npr, 5000):
        want = (b"k%07d" % keys[i] + b"x" * (keys[i] % 5)) if keys[i] < nb else None
        assert strs[i] == want
    lens = np.diff(got.cols[3].offsets)
    assert np.array_equal(lens, np.where(hit, 8 + keys % 5, 0))
