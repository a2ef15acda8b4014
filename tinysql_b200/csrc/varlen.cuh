// varlen.cuh — FLOAT (4-byte) and var-len (offsets + bytes) payload columns of HashJoinExec.
//
// The join kernels move 8-byte slots.  A column of another width never takes part in key comparison (keys are the
// 8-byte types), so it is kept in a side store on the device and represented inside the join by its ROW ID (an INT64
// column carrying the original NULL bitmap).  After the probe the row ids of the result batch are gathered back into a
// column of the original layout — the device form of chunk.CopySelectedJoinRows' per-cell copy
// (util/chunk/chunk_util.go:38-66: fixed cells copy elemLen bytes, var-len cells copy data[offsets[i]:offsets[i+1]]).
#pragma once
#include "common.cuh"

namespace tq {

struct SideStore {          // the cells of one column, device resident
  DevBuf offsets, bytes;    // var-len: int64 offsets[n+1] (relative to `base`) + bytes;  FLOAT: n 4-byte slots in `bytes`
  int64_t n = 0, base = 0;
  int elem = 0;             // 0 = var-len, 4 = FLOAT
};

struct VarOut {             // one gathered output column of a result batch
  DevBuf off, bytes;        // var-len: int64 off[n+1] starting at 0
  PinBuf h_off, h_bytes;
  int64_t total = 0;        // bytes in `bytes`
  int elem = 0;
  bool used = false, on_host = false;
};

int32_t iota_u64(uint64_t *dst, int64_t n, cudaStream_t s);
// out = cells store[rowids[i]] for i in [0, n); a 0 bit in bm makes cell i NULL (empty).  Synchronises s.
int32_t gather_cells(const SideStore &st, const uint64_t *rowids, const uint32_t *bm, int64_t n, VarOut &out, DevBuf &lens, DevBuf &scan_scratch,
                     cudaStream_t s);

}  // namespace tq
