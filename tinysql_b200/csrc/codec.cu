// codec.cu — the chunk wire format: chunk.Codec.Encode / Decode / DecodeToChunk
// (util/chunk/codec.go:42-143), the bytes child readers hand up (distsql/select_result.go:102-141).  Per column:
//     u32 length | u32 nullCount | nullBitmap[(length+7)/8] (only if nullCount > 0)
//     | int64 offsets[length+1] (var-len columns only) | data (length * fixedLen bytes, or offsets[length] bytes)
// Host-only code (no kernel, no device needed): decoding yields tq_column VIEWS into the buffer — zero copy — which can be
// passed straight to tq_join_put_probe / tq_agg_put; when the buffer lives in pinned memory (tq_pinned_alloc) the
// scan -> join path needs no host-side copy at all.
#include <cstring>

#include "common.cuh"

namespace {

int fixed_len(int32_t type) {  // getFixedLen, util/chunk/codec.go:171-181
  switch (type & 0xFF) {
    case TQ_TYPE_INT64: case TQ_TYPE_UINT64: case TQ_TYPE_FLOAT64: return 8;
    case TQ_TYPE_FLOAT32: return 4;
    case TQ_TYPE_BYTES: return -1;
  }
  return 0;
}

int64_t null_count(const tq_column &c) {  // Column.nullCount, column.go:94-104: zero bits among the first `length`
  if (!c.null_bitmap) return 0;
  int64_t ones = 0;
  const int64_t full = c.length >> 3;
  for (int64_t i = 0; i < full; i++) ones += __builtin_popcount(c.null_bitmap[i]);
  if (c.length & 7) ones += __builtin_popcount(c.null_bitmap[full] & ((1u << (c.length & 7)) - 1));
  return c.length - ones;
}

}  // namespace

extern "C" {

int32_t tq_chunk_encoded_size(int32_t n_cols, const int32_t *types, const tq_column *cols, int64_t *bytes) {
  if (n_cols < 0 || !bytes || (n_cols && (!types || !cols))) return TQ_ERR_INVALID_ARG;
  int64_t total = 0;
  for (int c = 0; c < n_cols; c++) {
    const int fl = fixed_len(types[c]);
    if (fl == 0 || cols[c].length < 0 || cols[c].length > 0xFFFFFFFFll) { tq::set_error("chunk codec: bad column %d", c); return TQ_ERR_UNSUPPORTED_TYPE; }
    const int64_t n = cols[c].length;
    total += 8;
    if (null_count(cols[c]) > 0) total += (n + 7) >> 3;
    if (fl < 0) {
      if (!cols[c].offsets) { tq::set_error("chunk codec: var-len column %d without offsets", c); return TQ_ERR_INVALID_ARG; }
      total += (n + 1) * 8 + (cols[c].offsets[n] - cols[c].offsets[0]);
    } else total += n * fl;
  }
  *bytes = total;
  return TQ_OK;
}

// Codec.Encode (codec.go:42-79).  Offsets are written rebased to 0, the way a freshly built Column holds them.
int32_t tq_chunk_encode(int32_t n_cols, const int32_t *types, const tq_column *cols, uint8_t *buffer, int64_t capacity, int64_t *written) {
  int64_t need = 0;
  TQ_TRY(tq_chunk_encoded_size(n_cols, types, cols, &need));
  if (!buffer || !written || capacity < need) { tq::set_error("chunk codec: buffer of %lld bytes needed", (long long)need); return TQ_ERR_INVALID_ARG; }
  uint8_t *p = buffer;
  for (int c = 0; c < n_cols; c++) {
    const tq_column &col = cols[c];
    const int64_t n = col.length;
    const uint32_t len32 = (uint32_t)n, nulls32 = (uint32_t)null_count(col);
    memcpy(p, &len32, 4); p += 4;
    memcpy(p, &nulls32, 4); p += 4;
    if (nulls32 > 0) {
      const int64_t nb = (n + 7) >> 3;
      memcpy(p, col.null_bitmap, (size_t)nb);
      if (n & 7) p[nb - 1] &= (uint8_t)((1u << (n & 7)) - 1);
      p += nb;
    }
    const int fl = fixed_len(types[c]);
    if (fl < 0) {
      const int64_t base = col.offsets[0];
      for (int64_t i = 0; i <= n; i++) { const int64_t o = col.offsets[i] - base; memcpy(p, &o, 8); p += 8; }
      const int64_t db = col.offsets[n] - base;
      if (db) memcpy(p, col.data + base, (size_t)db);
      p += db;
    } else {
      if (n) memcpy(p, col.data, (size_t)(n * fl));
      p += n * fl;
    }
  }
  *written = p - buffer;
  return TQ_OK;
}

// Codec.DecodeToChunk (codec.go:92-143): fills out[c] with views INTO buffer (null_bitmap == NULL when the column has
// no NULLs: the all-ones bitmap of setAllNotNull is implied).  The int64 offsets of var-len columns are used in place,
// so the buffer must keep them 8-byte aligned for consumers that require it (Go reinterprets them the same way,
// bytesToI64Slice).  *consumed = bytes used; the rest of the buffer belongs to the next chunk.
int32_t tq_chunk_decode(const uint8_t *buffer, int64_t len, int32_t n_cols, const int32_t *types, tq_column *out, int64_t *consumed) {
  if (!buffer || len < 0 || n_cols < 0 || (n_cols && (!types || !out)) || !consumed) return TQ_ERR_INVALID_ARG;
  const uint8_t *p = buffer, *end = buffer + len;
  for (int c = 0; c < n_cols; c++) {
    const int fl = fixed_len(types[c]);
    if (fl == 0) { tq::set_error("chunk codec: unsupported column type %d", types[c]); return TQ_ERR_UNSUPPORTED_TYPE; }
    if (end - p < 8) { tq::set_error("chunk codec: truncated header of column %d", c); return TQ_ERR_INVALID_ARG; }
    uint32_t len32, nulls32;
    memcpy(&len32, p, 4);
    memcpy(&nulls32, p + 4, 4);
    p += 8;
    const int64_t n = len32;
    out[c].length = n;
    out[c].null_bitmap = nullptr;
    out[c].offsets = nullptr;
    out[c].data = nullptr;
    if (nulls32 > 0) {
      const int64_t nb = (n + 7) >> 3;
      if (end - p < nb) { tq::set_error("chunk codec: truncated bitmap of column %d", c); return TQ_ERR_INVALID_ARG; }
      out[c].null_bitmap = const_cast<uint8_t *>(p);
      p += nb;
    }
    int64_t data_bytes = n * fl;
    if (fl < 0) {
      const int64_t ob = (n + 1) * 8;
      if (end - p < ob) { tq::set_error("chunk codec: truncated offsets of column %d", c); return TQ_ERR_INVALID_ARG; }
      out[c].offsets = reinterpret_cast<int64_t *>(const_cast<uint8_t *>(p));
      memcpy(&data_bytes, p + n * 8, 8);  // offsets[length]
      p += ob;
      if (data_bytes < 0) { tq::set_error("chunk codec: corrupt offsets of column %d", c); return TQ_ERR_INVALID_ARG; }
    }
    if (end - p < data_bytes) { tq::set_error("chunk codec: truncated data of column %d", c); return TQ_ERR_INVALID_ARG; }
    out[c].data = const_cast<uint8_t *>(p);
    p += data_bytes;
  }
  *consumed = p - buffer;
  return TQ_OK;
}

}  // extern "C"
