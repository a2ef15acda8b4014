"""GPU parity for the device-side chunk decoder (tq_chunk_decode_device / k_chunk_unpack): the wire bytes of
chunk.Codec.Encode (util/chunk/codec.go:42-79) are shipped as they are and unpacked in HBM; the columns must equal what
Codec.DecodeToChunk (:92-143) yields on the host — bit-exact, every layout, every alignment — and must be usable by the
operators' TQ_MEM_DEVICE entry points without touching the host again."""
import ctypes as C

import numpy as np
import pytest

import oracle_py as O
from test_chunk_codec import rand_cells
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, UINT64, Chunk, Column, DeviceChunk, decode_chunk, device_to_host, encode_chunk, tq_array
from tinysql_b200.executor import AGG_COUNT, AGG_FIRSTROW, AGG_MAX, AGG_SUM
from util import assert_same_multiset, gen_col

pytestmark = pytest.mark.gpu


def assert_chunks_equal(got, want):
    assert len(got.cols) == len(want.cols)
    for a, b in zip(got.cols, want.cols):
        assert a.length == b.length and a.tp == b.tp
        assert np.array_equal(a.not_null(), b.not_null())
        if a.tp == BYTES:
            assert a.tolist() == b.tolist()
        else:
            m = a.not_null()
            assert np.array_equal(a.raw()[m], b.raw()[m])


def test_reference_test_codec_golden_on_device(lib):
    # codec_test.go:29-69: 4 columns x 10 rows — NULL, int64(i), "i.12345", "i.12345"
    n = 10
    strs = [("%d.12345" % i).encode() for i in range(n)]
    types = [INT64, INT64, BYTES, BYTES]
    cols = [Column(INT64, np.zeros(n), [False] * n), Column(INT64, np.arange(n)), Column(BYTES, strs), Column(BYTES, strs)]
    buf = encode_chunk(types, cols)
    d = DeviceChunk().decode(buf, types)
    assert d.consumed == len(buf)
    chk = d.to_host()
    assert chk.cols[0].tolist() == [None] * n and chk.cols[1].tolist() == list(range(n))
    assert chk.cols[2].tolist() == strs and chk.cols[3].tolist() == strs
    d.free()


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 63, 64, 65, 1000, 100003])
@pytest.mark.parametrize("null_frac", [0.0, 0.3])
def test_device_decode_equals_host_decode(lib, n, null_frac):
    rng = np.random.default_rng(n * 5 + int(null_frac * 10))
    # the FLOAT and var-len columns in front shift every later column to an odd byte offset inside the blob
    types = [FLOAT32, BYTES, INT64, UINT64, FLOAT64, BYTES, FLOAT32, INT64]
    cols = [Column(FLOAT32, rng.random(n).astype(np.float32), (rng.random(n) >= null_frac) if null_frac else None),
            Column(BYTES, rand_cells(rng, n, null_frac)),
            gen_col(rng, INT64, n, null_frac), gen_col(rng, UINT64, n, null_frac), gen_col(rng, FLOAT64, n, null_frac),
            Column(BYTES, rand_cells(rng, n, null_frac)),
            Column(FLOAT32, rng.random(n).astype(np.float32)), gen_col(rng, INT64, n, null_frac)]
    buf = encode_chunk(types, cols)
    want, used = decode_chunk(buf, types)
    d = DeviceChunk()
    for _ in range(2):   # the second decode reuses the handle's device memory
        d.decode(buf + buf, types)    # two chunks back to back: `consumed` finds the boundary
        assert d.consumed == used == len(buf)
        assert_chunks_equal(d.to_host(), want)
        for v in d.tq_cols:   # the alignment the TQ_MEM_DEVICE entry points ask for (include/tinysql_b200.h: tq_column)
            assert v.data % 16 == 0 and (not v.null_bitmap or v.null_bitmap % 8 == 0) and (not v.offsets or v.offsets % 8 == 0)
    d.free()


def test_decoded_columns_feed_the_operators_in_hbm(lib):
    """scan -> join / agg without a host copy of the payload: device-decoded columns go straight into TQ_MEM_DEVICE puts"""
    rng = np.random.default_rng(77)
    nb, npr = 30000, 250001
    b = [Column(INT64, rng.permutation(nb)), gen_col(rng, INT64, nb, 0.0, 0, 1000)]
    p = [gen_col(rng, INT64, npr, 0.0, 0, nb + 5000), Column(INT64, np.arange(npr))]
    types = [INT64, INT64]
    db, dp = DeviceChunk().decode(encode_chunk(types, b), types), DeviceChunk().decode(encode_chunk(types, p), types)
    t, k = (C.c_int32 * 2)(1, 1), (C.c_int32 * 1)(0)
    d = L.TQJoinDesc(0, 1, 2, t, 2, t, 1, k, k, 0, 0)
    h = C.c_void_p()
    L.check(lib.tq_join_create(C.byref(d), C.byref(h)))
    L.check(lib.tq_join_put_build(h, db.tq_cols, L.TQ_MEM_DEVICE))
    L.check(lib.tq_join_finalize_build(h))
    L.check(lib.tq_join_put_probe(h, dp.tq_cols, None, L.TQ_MEM_DEVICE))
    L.check(lib.tq_join_probe_eof(h))
    out4 = (L.TQColumn * 4)()
    n, eof = C.c_int64(0), C.c_int32(0)
    L.check(lib.tq_join_next_device(h, out4, C.byref(n), C.byref(eof)))
    got = Chunk([device_to_host(INT64, out4[c].data, None, n.value) for c in range(4)])
    L.check(lib.tq_join_destroy(h))
    assert_same_multiset(got, O.hash_join(0, True, types, b, types, p, [0], [0]))
    # nullable columns: the unpacked bitmap words are what the aggregate kernels read
    g = gen_col(rng, INT64, npr, 0.1, 0, 700)
    x = gen_col(rng, FLOAT64, npr, 0.25)
    x.values[:] = np.floor(x.values)                      # exact sums in any order
    at = [INT64, FLOAT64]
    da = DeviceChunk().decode(encode_chunk(at, [g, x]), at)
    funcs = [(AGG_SUM, 1), (AGG_COUNT, 1), (AGG_MAX, 1), (AGG_COUNT, -1), (AGG_FIRSTROW, 0)]
    it, gb = (C.c_int32 * 2)(*at), (C.c_int32 * 1)(0)
    fa = (L.TQAggFunc * len(funcs))(*[L.TQAggFunc(f, a) for f, a in funcs])
    ad = L.TQAggDesc(2, it, 1, gb, len(funcs), fa, 700)
    h = C.c_void_p()
    L.check(lib.tq_agg_create(C.byref(ad), C.byref(h)))
    L.check(lib.tq_agg_put(h, da.tq_cols, L.TQ_MEM_DEVICE))
    L.check(lib.tq_agg_eof(h))
    out = [Column.empty(tp, 4096) for tp in (FLOAT64, INT64, FLOAT64, INT64, INT64)]
    L.check(lib.tq_agg_next(h, 4096, tq_array(out, 4096), C.byref(n), C.byref(eof)))
    L.check(lib.tq_agg_destroy(h))
    got = Chunk([Column(c.tp, c.values[: n.value], c.not_null()[: n.value]) for c in out])
    rc, want = O.hash_agg(at, [g, x], [0], funcs)
    assert rc == 0
    assert_same_multiset(got, want)
    for dc in (db, dp, da):
        dc.free()


def test_device_decode_rejects_truncated_buffers(lib):
    types = [INT64, BYTES]
    buf = encode_chunk(types, [Column(INT64, [1, 2, 3], [True, False, True]), Column(BYTES, [b"ab", None, b"cdef"])])
    d = DeviceChunk()
    for cut in (0, 4, 9, 20, len(buf) - 1):
        with pytest.raises(L.TQError):
            d.decode(buf[:cut], types)
    d.free()
