// dict.cuh — exact multi-column key encoding for HashJoin / HashAgg.
//
// The reference handles an N-column join key / GROUP BY list by hashing the concatenated encoded key bytes and then
// verifying candidates by value (hashRowContainer.matchJoinKey executor/hash_table.go:128-141 + codec.EqualChunkRow
// util/codec/codec.go:363-382; getGroupKey executor/aggregate.go:359-394 builds the exact encoded key bytes).  The
// device operators work on ONE 64-bit key word per row, so an N-column key is first mapped — exactly, never by a
// lossy hash — to one 64-bit word:
//   * every key column gets a dictionary value -> dense 32-bit id (ids 0 / 1 are reserved for NULL and for the
//     value that equals the table's empty marker);
//   * ids are folded left to right: pair = id_a << 32 | id_b; with more than two columns the pair itself is
//     dictionary-encoded back to 32 bits before the next column is folded in.
// Equal key tuples get equal words and different tuples different words by construction, so the single-key kernels
// keep the reference's equality semantics with no false matches.
#pragma once
#include "common.cuh"

namespace tq {

static constexpr int MK_MAX_KEYS = 8;
static constexpr uint32_t MK_ID_NULL = 0u, MK_ID_EMPTYVAL = 1u, MK_ID_MISS = 0xFFFFFFFFu;
static constexpr uint64_t MK_PAIR_MISS = 0xFFFFFFFFFFFFFFFFull;

// value -> id open-addressed dictionary (device resident, grows by rehash, load factor <= 0.5)
struct KeyDict {
  DevBuf keys, ids, meta;  // meta: u32 next regular id offset
  uint64_t n_slots = 0;
  uint64_t count = 0;      // regular values inserted so far (host mirror, exact after every insert)
  int32_t ensure(uint64_t extra, cudaStream_t s);
  // insert every non-NULL value of vals[0..n) (values equal to MK_PAIR_MISS are skipped when skip_miss is set)
  int32_t insert(const uint64_t *vals, const uint32_t *bm, int64_t n, bool skip_miss, cudaStream_t s);
  // out[r] = id of vals[r]; NULL -> null_id; absent -> MK_ID_MISS; no_signbit: values with bit 63 set never match
  int32_t lookup(const uint64_t *vals, const uint32_t *bm, int64_t n, uint32_t null_id, bool no_signbit, bool skip_miss, uint32_t *out,
                 cudaStream_t s) const;
};

struct MultiKeyEncoder {
  int k = 0;
  KeyDict col[MK_MAX_KEYS];
  KeyDict fold[MK_MAX_KEYS];
  DevBuf acc, tmp, pair;
  // Encode the k key columns of n rows into out_comb[0..n).
  //   insert          new values extend the dictionaries (build side / aggregation input); otherwise lookup only
  //   null_is_value   NULL is a key value of its own (GROUP BY); otherwise a NULL in any column invalidates the row
  //   no_signbit[i]   column i is compared across signedness: values with bit 63 set cannot match (join probe side)
  //   out_bm          optional NOT-NULL bitmap of the encoded column: bit = 0 for rows without a valid encoding
  int32_t encode(const DCol *keycols, const bool *no_signbit, int64_t n, bool insert, bool null_is_value, uint64_t *out_comb, uint32_t *out_bm,
                 cudaStream_t s);
  void release();
};

}  // namespace tq
