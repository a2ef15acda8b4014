"""util/chunk mirror: Column / Chunk laid out byte-for-byte like Go's chunk.Column
(util/chunk/column.go:28-34, Appendix A of SURVEY.md) on top of numpy buffers, plus the
pinned / device bridges the C-ABI works with.  Test-harness and benchmark plumbing only: the
operators themselves live in libtinysql_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib as L

MAX_CHUNK_SIZE = 1024  # DefMaxChunkSize sessionctx/variable/tidb_vars.go:241

INT64, UINT64, FLOAT64 = L.TQ_TYPE_INT64, L.TQ_TYPE_UINT64, L.TQ_TYPE_FLOAT64
_NP = {INT64: np.int64, UINT64: np.uint64, FLOAT64: np.float64}


def bitmap_bytes(n):
    return (n + 7) >> 3


def pack_not_null(not_null):
    """bool[n] -> nullBitmap bytes (bit i&7 of byte i>>3, 1 = NOT NULL; column.go:89-92)."""
    return np.packbits(np.asarray(not_null, dtype=bool), bitorder="little")


def unpack_not_null(bitmap, n):
    if bitmap is None:
        return np.ones(n, dtype=bool)
    return np.unpackbits(np.asarray(bitmap, dtype=np.uint8), count=n, bitorder="little").astype(bool)


class Column:
    """One fixed-width 8-byte chunk.Column in host memory."""

    def __init__(self, tp, values, not_null=None):
        self.tp = tp
        v = np.ascontiguousarray(values)
        if v.dtype != _NP[tp]:
            v = v.astype(_NP[tp])
        self.values = v
        self.length = int(v.shape[0])
        self.bitmap = None if not_null is None else pack_not_null(not_null)

    @classmethod
    def empty(cls, tp, n):
        """Caller-allocated result column (data n*8 bytes + ceil(n/8) bitmap bytes)."""
        c = cls(tp, np.zeros(n, dtype=_NP[tp]))
        c.bitmap = np.zeros(bitmap_bytes(max(n, 1)), dtype=np.uint8)
        return c

    def not_null(self):
        return unpack_not_null(self.bitmap, self.length)

    def raw(self):
        return self.values.view(np.uint64)

    def slice(self, lo, hi):
        nn = None if self.bitmap is None else self.not_null()[lo:hi]
        return Column(self.tp, self.values[lo:hi].copy(), nn)

    def tq(self, length=None):
        t = L.TQColumn()
        t.length = self.length if length is None else length
        t.data = self.values.ctypes.data if self.values.size else None
        t.null_bitmap = self.bitmap.ctypes.data if self.bitmap is not None else None
        t.offsets = None
        return t

    def tolist(self):
        nn = self.not_null()
        return [self.values[i].item() if nn[i] else None for i in range(self.length)]


def tq_array(cols, length=None):
    arr = (L.TQColumn * max(len(cols), 1))()
    for i, c in enumerate(cols):
        arr[i] = c.tq(length)
    return arr


class Chunk:
    """chunk.Chunk: a list of equally long columns (no sel vector on the operator boundary)."""

    def __init__(self, cols):
        self.cols = list(cols)

    def num_rows(self):
        return self.cols[0].length if self.cols else 0

    @classmethod
    def concat(cls, chunks, types):
        cols = []
        for ci, tp in enumerate(types):
            vals = [c.cols[ci].values for c in chunks if c.num_rows()]
            nns = [c.cols[ci].not_null() for c in chunks if c.num_rows()]
            if vals:
                cols.append(Column(tp, np.concatenate(vals), np.concatenate(nns)))
            else:
                cols.append(Column(tp, np.zeros(0, dtype=_NP[tp]), np.zeros(0, dtype=bool)))
        return cls(cols)

    def rows(self):
        """list of tuples with None for NULL — the `testkit.Rows` view of a result."""
        lists = [c.tolist() for c in self.cols]
        return list(zip(*lists)) if lists else []


class DeviceColumn:
    """A column in HBM allocated through the C-ABI (tq_device_alloc)."""

    def __init__(self, tp, n, with_bitmap=True):
        lib = L.load()
        self.tp, self.length = tp, int(n)
        self._data = C.c_void_p()
        L.check(lib.tq_device_alloc(max(self.length, 1) * 8, C.byref(self._data)))
        self._bm = C.c_void_p()
        self.bm_bytes = (((self.length + 63) >> 6) << 3) + 8  # 8-byte padded: kernels use 32-bit words
        if with_bitmap:
            L.check(lib.tq_device_alloc(self.bm_bytes, C.byref(self._bm)))
            L.check(lib.tq_memset_device(self._bm, 0, self.bm_bytes))

    @classmethod
    def from_host(cls, col):
        d = cls(col.tp, col.length, with_bitmap=col.bitmap is not None)
        lib = L.load()
        if col.length:
            L.check(lib.tq_memcpy_h2d(d._data, col.values.ctypes.data, col.length * 8))
            if col.bitmap is not None:
                L.check(lib.tq_memcpy_h2d(d._bm, col.bitmap.ctypes.data, bitmap_bytes(col.length)))
        return d

    def tq(self, length=None):
        t = L.TQColumn()
        t.length = self.length if length is None else length
        t.data = self._data.value
        t.null_bitmap = self._bm.value
        t.offsets = None
        return t

    def to_host(self, n=None):
        n = self.length if n is None else n
        return device_to_host(self.tp, self._data.value, self._bm.value, n)

    def free(self):
        lib = L.load()
        if self._data:
            lib.tq_device_free(self._data)
            self._data = C.c_void_p()
        if self._bm:
            lib.tq_device_free(self._bm)
            self._bm = C.c_void_p()


def device_to_host(tp, data_ptr, bm_ptr, n):
    lib = L.load()
    out = Column.empty(tp, n)
    if n:
        L.check(lib.tq_memcpy_d2h(out.values.ctypes.data, data_ptr, n * 8))
        if bm_ptr:
            L.check(lib.tq_memcpy_d2h(out.bitmap.ctypes.data, bm_ptr, bitmap_bytes(n)))
        else:
            out.bitmap = None
    return out
