// strdict.cuh — exact byte-string -> dense id dictionary on the device: var-len (VARCHAR / BLOB ...) join keys, GROUP BY
// items and string aggregate arguments.
//
// Reference semantics: a string key value is encoded as compactBytesFlag + its raw bytes (util/codec/codec.go:226-233,
// 318-333) and two key values are equal iff flag and bytes are equal (EqualChunkRow, codec.go:363-382); a string GROUP BY
// item contributes encodeBytes(bytes) to the group key (HashGroupKey ETString, codec.go:735-743), which is injective on the
// bytes.  So "same string" is plain byte equality — no collation, no padding.  The device operators work on one 64-bit key
// word per row, therefore every distinct string gets a dense id (its index in an append-only arena of the distinct
// strings) and the id is the key.  Equality is decided by comparing BYTES, never by the hash alone: two different
// strings with equal 64-bit hashes still get different ids.
#pragma once
#include "common.cuh"
#include "varlen.cuh"

namespace tq {

static constexpr uint64_t SD_ID_MISS = 0xFFFFFFFFFFFFFFFFull;  // lookup of a string the dictionary does not hold

// cells of a var-len column as they sit in a SideStore: cell r = bytes[off[r] - base, off[r + 1] - base)
struct StrView {
  const int64_t *off = nullptr;
  int64_t base = 0;
  const uint8_t *bytes = nullptr;
};
inline StrView view_of(const SideStore &s) {
  StrView v;
  v.off = s.offsets.as<int64_t>();
  v.base = s.base;
  v.bytes = s.bytes.as<uint8_t>();
  return v;
}

struct StringDict {
  SideStore arena;          // the distinct strings: offsets[count + 1] (base 0) + bytes — id i = cell i (gather_cells reads it)
  DevBuf a_hash;            // u64 hash of string i
  DevBuf slots;             // u32 per slot: SD_EMPTY, an arena id, or (during an insert) SD_BATCH | batch row
  uint64_t n_slots = 0;
  uint64_t count = 0;       // distinct strings
  uint64_t used_bytes = 0;  // bytes of the arena in use
  size_t off_cap = 0, hash_cap = 0, bytes_cap = 0;  // valid capacities of the arena arrays (entries / bytes)
  DevBuf row_hash, new_src, new_len, new_off, meta, scan_scratch;

  // ids_out[r] = id of cell r (as u64).  valid_out (optional, 32-bit words, ((n + 31) / 32) words written): bit r = the cell is
  // NOT NULL and — lookup mode — present in the dictionary; ids of invalid rows are SD_ID_MISS.
  //   insert = true : every non-NULL cell is added first (build side / aggregation input)
  //   batch_bytes   : total bytes of the n cells (bounds the arena growth of an insert)
  int32_t encode(const StrView &v, const uint32_t *bm, int64_t n, int64_t batch_bytes, bool insert, uint64_t *ids_out, uint32_t *valid_out, cudaStream_t s);
  // MAX / MIN over string arguments: three-way byte-wise compare of two arena strings is done on the device (see agg.cu)
  void release();

 private:
  int32_t ensure(uint64_t extra_strings, uint64_t extra_bytes, cudaStream_t s);
};

#ifdef __CUDACC__
// types.CompareString (types/compare.go:115-123) = bytes.Compare of the two cells; one thread.
__device__ __forceinline__ int sd_compare_ids(const int64_t *a_off, const uint8_t *a_bytes, uint32_t x, uint32_t y) {
  const int64_t x0 = a_off[x], xl = a_off[x + 1] - x0, y0 = a_off[y], yl = a_off[y + 1] - y0;
  const int64_t m = xl < yl ? xl : yl;
  for (int64_t i = 0; i < m; i++) {
    const uint8_t a = a_bytes[x0 + i], b = a_bytes[y0 + i];
    if (a != b) return a < b ? -1 : 1;
  }
  return xl < yl ? -1 : (xl == yl ? 0 : 1);
}
#endif

}  // namespace tq
