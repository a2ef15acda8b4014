"""CPU suite: the chunk wire codec of the C-ABI (tq_chunk_encode / tq_chunk_decode, host-only code) against an independent
Python restatement of chunk.Codec (util/chunk/codec.go:42-143) and the reference's TestCodec (codec_test.go:29-69)."""
import struct

import numpy as np
import pytest

from tinysql_b200 import _lib as L
from tinysql_b200.chunk import BYTES, FLOAT32, FLOAT64, INT64, UINT64, Column, decode_chunk, encode_chunk
from util import gen_col


def py_encode(types, cols):
    """Codec.encodeColumn, statement by statement (codec.go:51-79)"""
    out = b""
    for tp, c in zip(types, cols):
        n = c.length
        nn = c.not_null()
        nulls = int(n - nn.sum())
        out += struct.pack("<II", n, nulls)
        if nulls > 0:
            out += np.packbits(nn, bitorder="little").tobytes()
        if tp == BYTES:
            out += c.offsets.astype("<i8").tobytes() + c.data[: c.offsets[n]].tobytes()
        else:
            out += c.values.tobytes()
    return out


def rand_cells(rng, n, null_frac):
    return [None if rng.random() < null_frac else rng.integers(0, 256, int(rng.integers(0, 20)), dtype=np.uint8).tobytes() for _ in range(n)]


def test_reference_test_codec_golden():
    # codec_test.go:29-69: 4 columns x 10 rows — NULL, int64(i), "i.12345", "i.12345"
    n = 10
    strs = [("%d.12345" % i).encode() for i in range(n)]
    types = [INT64, INT64, BYTES, BYTES]
    cols = [Column(INT64, np.zeros(n), [False] * n), Column(INT64, np.arange(n)), Column(BYTES, strs), Column(BYTES, strs)]
    buf = encode_chunk(types, cols)
    assert buf == py_encode(types, cols)
    chk, used = decode_chunk(buf, types)
    assert used == len(buf)                      # c.Assert(len(remained), check.Equals, 0)
    assert len(chk.cols) == 4 and chk.num_rows() == n
    assert chk.cols[0].tolist() == [None] * n    # row.IsNull(0)
    assert chk.cols[1].tolist() == list(range(n))
    assert chk.cols[2].tolist() == strs and chk.cols[3].tolist() == strs


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1000])
@pytest.mark.parametrize("null_frac", [0.0, 0.3])
def test_codec_round_trip_and_bytes_match_python_restatement(n, null_frac):
    rng = np.random.default_rng(n * 3 + int(null_frac * 10))
    types = [INT64, UINT64, FLOAT64, FLOAT32, BYTES]
    cols = [gen_col(rng, INT64, n, null_frac), gen_col(rng, UINT64, n, null_frac), gen_col(rng, FLOAT64, n, null_frac),
            Column(FLOAT32, rng.random(n).astype(np.float32), (rng.random(n) >= null_frac) if null_frac else None),
            Column(BYTES, rand_cells(rng, n, null_frac))]
    buf = encode_chunk(types, cols)
    assert buf == py_encode(types, cols)
    # two chunks back to back: `consumed` finds the boundary
    chk, used = decode_chunk(buf + buf, types)
    assert used == len(buf)
    for a, b in zip(chk.cols, cols):
        assert a.length == b.length
        assert np.array_equal(a.not_null(), b.not_null())
        m = a.not_null()
        if a.tp == BYTES:
            assert a.tolist() == b.tolist()
        else:
            assert np.array_equal(a.raw()[m], b.raw()[m])


def test_codec_rejects_truncated_buffers():
    types = [INT64, BYTES]
    cols = [Column(INT64, [1, 2, 3], [True, False, True]), Column(BYTES, [b"ab", None, b"cdef"])]
    buf = encode_chunk(types, cols)
    for cut in (0, 4, 9, 20, len(buf) - 1):
        with pytest.raises(L.TQError):
            decode_chunk(buf[:cut], types)
