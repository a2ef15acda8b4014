// codec.cu — the chunk wire format: chunk.Codec.Encode / Decode / DecodeToChunk
// (util/chunk/codec.go:42-143), the bytes child readers hand up (distsql/select_result.go:102-141).  Per column:
//     u32 length | u32 nullCount | nullBitmap[(length+7)/8] (only if nullCount > 0)
//     | int64 offsets[length+1] (var-len columns only) | data (length * fixedLen bytes, or offsets[length] bytes)
// Host side (no kernel, no device needed): decoding yields tq_column VIEWS into the buffer — zero copy — which can be
// passed straight to tq_join_put_probe / tq_agg_put; when the buffer lives in pinned memory (tq_pinned_alloc) the
// scan -> join path needs no host-side copy at all.
// Device side (tq_chunk_decode_device): the wire bytes cross PCIe ONCE, as they are, and k_chunk_unpack lays the columns
// out in HBM (8-byte aligned data / offsets / bitmap words, tail bits zeroed) ready for the TQ_MEM_DEVICE entry points —
// the host never touches the payload.
#include <cstring>
#include <vector>

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

// One contiguous piece of a wire-format column (bitmap, offsets or data) on its way to an aligned device array.
struct UnpackSeg {
  uint64_t src_off;     // byte offset inside the blob — ANY alignment (the wire format packs columns back to back)
  uint64_t *dst;        // 8-byte aligned destination
  uint64_t n_words;     // 8-byte words to write (the destination is padded to whole words)
  uint64_t valid_bits;  // bits of the piece that carry data; everything above is written as 0
};
static constexpr int UNPACK_MAX_SEGS = 48;
struct UnpackParams {
  const uint8_t *blob;
  int n_segs;
  UnpackSeg seg[UNPACK_MAX_SEGS];
};

// blockIdx.y = piece; grid-stride over its destination words.  A destination word is assembled from the two aligned
// source words it straddles (funnel shift by the piece's misalignment); loads stay inside [blob, blob + len + 16).
__global__ void __launch_bounds__(256) k_chunk_unpack(const UnpackParams p) {
  const UnpackSeg s = p.seg[blockIdx.y];
  const uint64_t *a = reinterpret_cast<const uint64_t *>(p.blob + (s.src_off & ~7ull));
  const unsigned sh = (unsigned)(s.src_off & 7ull) * 8u;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n_words; i += stride) {
    uint64_t v = 0;
    const uint64_t first_bit = i * 64;
    if (first_bit < s.valid_bits) {
      const uint64_t lo = a[i];
      v = lo;
      if (sh) v = (lo >> sh) | (a[i + 1] << (64u - sh));
      const uint64_t keep = s.valid_bits - first_bit;
      if (keep < 64) v &= (1ull << keep) - 1ull;
    }
    s.dst[i] = v;
  }
}

}  // namespace

struct tq_chunk_device {
  tq::DevBuf blob, arena;   // grow-only: a handle that is reused decodes without allocating
};

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

// Decoder for the device (util/chunk/codec.go:92-143, 246-353; the consumer side of distsql/select_result.go:102-141).
// Parses the headers on the host (a few bytes per column), ships buffer[0, consumed) to HBM in one copy and unpacks every
// column with ONE kernel launch.  out[c] receives DEVICE pointers (null_bitmap == NULL: no NULLs; var-len columns:
// offsets + data; FLOAT: 4-byte slots), valid until the handle is decoded into again or freed.
int32_t tq_chunk_decode_device(const uint8_t *buffer, int64_t len, int32_t n_cols, const int32_t *types, tq_chunk_device **chunk,
                               tq_column *out, int64_t *consumed) {
  if (!chunk || (n_cols > 0 && !out) || n_cols > UNPACK_MAX_SEGS / 3) { if (chunk) tq::set_error("chunk codec: more than %d columns", UNPACK_MAX_SEGS / 3); return TQ_ERR_INVALID_ARG; }
  std::vector<tq_column> view((size_t)(n_cols > 0 ? n_cols : 1));
  TQ_TRY(tq_chunk_decode(buffer, len, n_cols, types, view.data(), consumed));   // validates sizes; views are host pointers
  TQ_TRY(tq::ensure_init());
  tq::Runtime &r = tq::rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  tq_chunk_device *h = *chunk ? *chunk : new tq_chunk_device();
  auto fail = [&](int32_t st) { if (!*chunk) delete h; return st; };
  // arena layout: per column bitmap words (+1 pad word: kernels read bitmaps as 32-bit words), offsets, data; each piece 16-byte aligned
  UnpackParams p{};
  uint64_t arena_words = 0;
  std::vector<uint64_t> dst_word((size_t)n_cols * 3 + 1, 0);
  for (int c = 0; c < n_cols; c++) {
    const tq_column &v = view[c];
    const int64_t n = v.length;
    const int fl = fixed_len(types[c]);
    if (v.null_bitmap) {
      UnpackSeg &sg = p.seg[p.n_segs];
      sg.src_off = (uint64_t)(v.null_bitmap - buffer);
      sg.valid_bits = (uint64_t)n;
      sg.n_words = (uint64_t)((n + 63) >> 6) + 1;
      dst_word[p.n_segs++] = arena_words;
      arena_words += (sg.n_words + 1) & ~1ull;   // every piece starts 16-byte aligned (the tq_vec_* / tq_expr_eval device rule)
    }
    if (fl < 0) {
      UnpackSeg &sg = p.seg[p.n_segs];
      sg.src_off = (uint64_t)(reinterpret_cast<const uint8_t *>(v.offsets) - buffer);
      sg.valid_bits = (uint64_t)(n + 1) * 64;
      sg.n_words = (uint64_t)(n + 1);
      dst_word[p.n_segs++] = arena_words;
      arena_words += (sg.n_words + 1) & ~1ull;   // every piece starts 16-byte aligned (the tq_vec_* / tq_expr_eval device rule)
    }
    int64_t data_bytes = n * fl;
    if (fl < 0) memcpy(&data_bytes, reinterpret_cast<const uint8_t *>(v.offsets) + n * 8, 8);
    {
      UnpackSeg &sg = p.seg[p.n_segs];
      sg.src_off = (uint64_t)(v.data - buffer);
      sg.valid_bits = (uint64_t)data_bytes * 8;
      sg.n_words = (uint64_t)((data_bytes + 7) >> 3) + 1;   // at least one word: data is never a NULL pointer
      dst_word[p.n_segs++] = arena_words;
      arena_words += (sg.n_words + 1) & ~1ull;   // every piece starts 16-byte aligned (the tq_vec_* / tq_expr_eval device rule)
    }
  }
  int32_t st = h->blob.reserve((size_t)*consumed + 32);
  if (st == TQ_OK) st = h->arena.reserve((size_t)(arena_words ? arena_words : 1) * 8);
  if (st != TQ_OK) return fail(st);
  uint64_t max_words = 0;
  for (int i = 0; i < p.n_segs; i++) { p.seg[i].dst = h->arena.as<uint64_t>() + dst_word[i]; if (p.seg[i].n_words > max_words) max_words = p.seg[i].n_words; }
  p.blob = h->blob.as<uint8_t>();
  cudaStream_t s = r.compute;
  if (*consumed > 0) {
    cudaError_t e = cudaMemcpyAsync(h->blob.p, buffer, (size_t)*consumed, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return fail(tq::cuda_fail(e, "cudaMemcpyAsync(chunk blob)", __FILE__, __LINE__));
  }
  if (p.n_segs > 0) {
    uint64_t bx = (max_words + 255) / 256;
    const uint64_t cap = (uint64_t)r.sm_count * 8;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    TQ_LAUNCH(k_chunk_unpack, dim3((unsigned)bx, (unsigned)p.n_segs), 256, 0, s, p);
    tq::count_launch();
    st = tq::check_launch("k_chunk_unpack");
    if (st != TQ_OK) return fail(st);
  }
  // the caller may reuse `buffer` and hand the columns to any entry point as soon as this returns
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(tq::cuda_fail(e, "cudaStreamSynchronize(chunk decode)", __FILE__, __LINE__));
  int sg = 0;
  for (int c = 0; c < n_cols; c++) {
    const tq_column &v = view[c];
    out[c].length = v.length;
    out[c].null_bitmap = nullptr;
    out[c].offsets = nullptr;
    if (v.null_bitmap) out[c].null_bitmap = reinterpret_cast<uint8_t *>(p.seg[sg++].dst);
    if (fixed_len(types[c]) < 0) out[c].offsets = reinterpret_cast<int64_t *>(p.seg[sg++].dst);
    out[c].data = reinterpret_cast<uint8_t *>(p.seg[sg++].dst);
  }
  *chunk = h;
  return TQ_OK;
}

int32_t tq_chunk_device_free(tq_chunk_device *chunk) {
  if (!chunk) return TQ_OK;
  tq::Runtime &r = tq::rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  delete chunk;
  return TQ_OK;
}

}  // extern "C"
