// strexpr.cu — vectorized builtins over var-len (string) columns.
//   builtin{LT,LE,GT,GE,EQ,NE}StringSig.vecEvalInt   expression/builtin_compare_vec_generated.go:65-555
//   builtinStrcmpSig.vecEvalInt                        expression/builtin_string_vec.go:52-83
//   builtinLengthSig                                   row version expression/builtin_string.go:75-81 (the vector
//                                                      version, builtin_string_vec.go:93-96, is a course stub)
//   builtinStringIsNullSig.vecEvalInt                  expression/builtin_string_vec.go:21-42
// Comparison is types.CompareString (types/compare.go:115-123): Go string order == unsigned byte-wise order, the
// shorter string first on a common prefix.  Result NULL iff either argument is NULL (MergeNulls), value 0 there.
//
// Byte streaming: one warp walks one row at a time with coalesced 32-byte reads and finds the first differing byte with a
// ballot; results of 32 consecutive rows are written as one coalesced store + one bitmap word.  HBM-bound.
#include "common.cuh"

namespace tq {

struct StrCol {
  const int64_t *off;   // n+1 offsets (any base)
  const uint8_t *data;  // cell i = data[off[i] - base .. off[i+1] - base)
  int64_t base;
  const uint32_t *bm;
};

__global__ void __launch_bounds__(256) k_str_compare(StrCol a, StrCol b, int op, int64_t n, uint64_t *out, uint32_t *out_bm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups = (n + 31) >> 5;
  for (int64_t g = warp; g < groups; g += n_warps) {
    const int64_t r0 = g << 5;
    long long my_val = 0;
    bool my_nn = false;
    for (int jrow = 0; jrow < 32; jrow++) {
      const int64_t r = r0 + jrow;
      if (r >= n) break;  // warp-uniform
      const bool nn = tqd::bm_not_null(a.bm, r) && tqd::bm_not_null(b.bm, r);
      long long val = 0;
      if (nn) {
        const int64_t a0 = a.off[r], a1 = a.off[r + 1], b0 = b.off[r], b1 = b.off[r + 1];
        const int64_t la = a1 - a0, lb = b1 - b0, m = la < lb ? la : lb;
        const uint8_t *pa = a.data + (a0 - a.base), *pb = b.data + (b0 - b.base);
        int cmp = 0;
        bool found = false;
        for (int64_t base = 0; base < m && !found; base += 32) {
          const int64_t i = base + lane;
          uint8_t ca = 0, cb = 0;
          if (i < m) { ca = pa[i]; cb = pb[i]; }
          const unsigned diff = __ballot_sync(0xffffffffu, ca != cb);
          if (diff) {
            const int src = __ffs(diff) - 1;
            const int xa = __shfl_sync(0xffffffffu, (int)ca, src), xb = __shfl_sync(0xffffffffu, (int)cb, src);
            cmp = xa < xb ? -1 : 1;
            found = true;
          }
        }
        if (!found) cmp = la < lb ? -1 : (la > lb ? 1 : 0);
        switch (op) {
          case 0: val = cmp < 0; break;    // LT
          case 1: val = cmp <= 0; break;   // LE
          case 2: val = cmp > 0; break;    // GT
          case 3: val = cmp >= 0; break;   // GE
          case 4: val = cmp == 0; break;   // EQ
          case 5: val = cmp != 0; break;   // NE
          default: val = cmp; break;       // STRCMP
        }
      }
      if (lane == jrow) { my_val = val; my_nn = nn; }
    }
    const int64_t r = r0 + lane;
    const unsigned word = __ballot_sync(0xffffffffu, my_nn);
    if (r < n) out[r] = (uint64_t)my_val;
    if (lane == 0) out_bm[g] = word;
  }
}

// op 0: LENGTH (bytes; NULL stays NULL)   op 1: IS NULL (never NULL)
__global__ void __launch_bounds__(256) k_str_unary(StrCol a, int op, int64_t n, uint64_t *out, uint32_t *out_bm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n + 31) & ~31ll;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += stride) {
    bool nn_out = false;
    if (r < n) {
      const bool nn = tqd::bm_not_null(a.bm, r);
      if (op == 0) { out[r] = nn ? (uint64_t)(a.off[r + 1] - a.off[r]) : 0ull; nn_out = nn; }
      else { out[r] = nn ? 0ull : 1ull; nn_out = true; }
    }
    const unsigned word = __ballot_sync(0xffffffffu, nn_out);
    if ((threadIdx.x & 31) == 0) out_bm[r >> 5] = word;
  }
}

// warp-cooperative byte equality of two cells (all 32 lanes call)
__device__ __forceinline__ bool str_cells_equal(const StrCol &a, const StrCol &b, int64_t r) {
  const int lane = threadIdx.x & 31;
  const int64_t a0 = a.off[r], la = a.off[r + 1] - a0, b0 = b.off[r], lb = b.off[r + 1] - b0;
  if (la != lb) return false;
  const uint8_t *pa = a.data + (a0 - a.base), *pb = b.data + (b0 - b.base);
  for (int64_t base = 0; base < la; base += 32) {
    const int64_t i = base + lane;
    const bool diff = i < la && pa[i] != pb[i];
    if (__any_sync(0xffffffffu, diff)) return false;
  }
  return true;
}

// builtinInStringSig.vecEvalInt (expression/builtin_other_vec_generated.go:97-149): 1 if some list element equals a
// (types.CompareString == 0); else NULL if a or any list element was NULL; else 0.
static constexpr int STR_MAX_IN_LIST = 8;
struct StrInList {
  int n;
  StrCol c[STR_MAX_IN_LIST];
};
__global__ void __launch_bounds__(256) k_str_in(StrCol a, StrInList L, int64_t n, uint64_t *out, uint32_t *out_bm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups = (n + 31) >> 5;
  for (int64_t g = warp; g < groups; g += n_warps) {
    const int64_t r0 = g << 5;
    long long my_val = 0;
    bool my_nn = false;
    for (int jrow = 0; jrow < 32; jrow++) {
      const int64_t r = r0 + jrow;
      if (r >= n) break;  // warp-uniform
      const bool ann = tqd::bm_not_null(a.bm, r);
      bool has_null = false, found = false;
      for (int j = 0; j < L.n; j++) {
        if (!ann || !tqd::bm_not_null(L.c[j].bm, r)) { has_null = true; continue; }
        if (!found && str_cells_equal(a, L.c[j], r)) found = true;
      }
      if (lane == jrow) { my_val = found ? 1 : 0; my_nn = found || !has_null; }
    }
    const int64_t r = r0 + lane;
    const unsigned word = __ballot_sync(0xffffffffu, my_nn);
    if (r < n) out[r] = (uint64_t)my_val;
    if (lane == 0) out_bm[g] = word;
  }
}

// String-valued IF / IFNULL: every row picks one of two source cells or NULL (builtinIfStringSig / builtinIfNullStringSig,
// expression/builtin_control_vec_generated.go:209-262, 81-112): lengths -> exclusive scan -> one warp copies one cell.
//   mode 0 (IF):      cond NULL or 0 -> b, else a          mode 1 (IFNULL): a unless NULL, then b
__global__ void __launch_bounds__(256) k_str_pick(int mode, const uint64_t *cond, const uint32_t *cond_bm, StrCol a, StrCol b, int64_t n, uint32_t *lens, uint8_t *src,
                                                   uint32_t *out_bm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (n + 31) & ~31ll;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += stride) {
    bool nn = false;
    if (r < n) {
      int pick;  // 1 = a, 2 = b
      if (mode == 0) pick = (tqd::bm_not_null(cond_bm, r) && cond[r] != 0) ? 1 : 2;
      else pick = tqd::bm_not_null(a.bm, r) ? 1 : 2;
      const StrCol &c = pick == 1 ? a : b;
      nn = tqd::bm_not_null(c.bm, r);
      lens[r] = nn ? (uint32_t)(c.off[r + 1] - c.off[r]) : 0u;
      src[r] = nn ? (uint8_t)pick : (uint8_t)0;
    }
    const unsigned word = __ballot_sync(0xffffffffu, nn);
    if ((threadIdx.x & 31) == 0) out_bm[r >> 5] = word;
  }
}
__global__ void __launch_bounds__(256) k_str_pick_copy(StrCol a, StrCol b, const uint8_t *src, const uint32_t *out_off32, const uint64_t *d_total, int64_t n,
                                                        int64_t *out_off, uint8_t *out_bytes) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (warp == 0 && lane == 0) out_off[n] = (int64_t)*d_total;
  for (int64_t r = warp; r < n; r += n_warps) {
    const uint32_t o = out_off32[r];
    if (lane == 0) out_off[r] = (int64_t)o;
    const int sr = src[r];
    if (sr == 0) continue;
    const StrCol &c = sr == 1 ? a : b;
    const int64_t s0 = c.off[r], len = c.off[r + 1] - s0;
    const uint8_t *p = c.data + (s0 - c.base);
    for (int64_t i = lane; i < len; i += 32) out_bytes[o + i] = p[i];
  }
}

namespace {

struct StrDev {  // device image of one var-len argument (owned when uploaded from the host)
  DevBuf off, data, bm;
  StrCol view{};
};

int32_t check_str_col(const tq_column *c, int64_t n, const char *what) {
  if (!c || (n > 0 && !c->offsets)) { set_error("%s: a var-len column (offsets + data) is required", what); return TQ_ERR_INVALID_ARG; }
  return TQ_OK;
}

int32_t stage_str(const tq_column *c, int64_t n, int32_t mem, StrDev &d, cudaStream_t s) {
  if (mem == TQ_MEM_DEVICE) {
    d.view.off = c->offsets;
    d.view.data = c->data;
    d.view.base = 0;  // device columns start at offset 0 by contract
    d.view.bm = (const uint32_t *)c->null_bitmap;
    return TQ_OK;
  }
  const int64_t b0 = n ? c->offsets[0] : 0, b1 = n ? c->offsets[n] : 0;
  if (b1 < b0 || (b1 > b0 && !c->data)) { set_error("malformed var-len column"); return TQ_ERR_INVALID_ARG; }
  TQ_TRY(d.off.reserve((size_t)(n + 1) * 8));
  TQ_TRY(d.data.reserve((size_t)(b1 - b0) + 16));
  if (n) TQ_CUDA(cudaMemcpyAsync(d.off.p, c->offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s));
  if (b1 > b0) TQ_CUDA(cudaMemcpyAsync(d.data.p, c->data + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, s));
  d.view.off = d.off.as<int64_t>();
  d.view.data = d.data.as<uint8_t>();
  d.view.base = b0;
  d.view.bm = nullptr;
  if (c->null_bitmap) {
    TQ_TRY(d.bm.reserve(bitmap_alloc_bytes(n)));
    TQ_CUDA(cudaMemcpyAsync(d.bm.p, c->null_bitmap, bitmap_bytes(n), cudaMemcpyHostToDevice, s));
    d.view.bm = d.bm.as<uint32_t>();
  }
  return TQ_OK;
}

template <typename Launch>
int32_t run_str(int64_t n, tq_column *out, int32_t mem, Launch &&launch) {
  if (!out || (n > 0 && (!out->data || !out->null_bitmap))) { set_error("output column needs data and null_bitmap buffers"); return TQ_ERR_INVALID_ARG; }
  Runtime &r = rt();
  cudaStream_t s = r.compute;
  out->length = n;
  if (n == 0) return TQ_OK;
  if (mem == TQ_MEM_DEVICE) {
    TQ_TRY(launch((uint64_t *)out->data, (uint32_t *)out->null_bitmap));
    TQ_CUDA(cudaStreamSynchronize(s));
    return TQ_OK;
  }
  DevBuf od, ob;
  TQ_TRY(od.reserve((size_t)n * 8));
  TQ_TRY(ob.reserve(bitmap_alloc_bytes(n)));
  TQ_TRY(launch(od.as<uint64_t>(), ob.as<uint32_t>()));
  TQ_CUDA(cudaMemcpyAsync(out->data, od.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaMemcpyAsync(out->null_bitmap, ob.p, bitmap_bytes(n), cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaStreamSynchronize(s));
  return TQ_OK;
}

int str_grid(int64_t threads) {
  int64_t blocks = (threads + 255) / 256;
  const int64_t cap = (int64_t)rt().sm_count * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace
}  // namespace tq

using namespace tq;

extern "C" {

int32_t tq_vec_compare_string(int32_t op, int64_t n, const tq_column *a, const tq_column *b, tq_column *out, int32_t mem) {
  TQ_TRY(ensure_init());
  if (op < TQ_CMP_LT || op > TQ_STR_STRCMP || n < 0) { set_error("bad string compare op %d", op); return TQ_ERR_INVALID_ARG; }
  TQ_TRY(check_str_col(a, n, "string compare lhs"));
  TQ_TRY(check_str_col(b, n, "string compare rhs"));
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  StrDev da, db;
  if (n) { TQ_TRY(stage_str(a, n, mem, da, r.compute)); TQ_TRY(stage_str(b, n, mem, db, r.compute)); }
  return run_str(n, out, mem, [&](uint64_t *od, uint32_t *ob) -> int32_t {
    k_str_compare<<<str_grid(((n + 31) >> 5) * 32), 256, 0, r.compute>>>(da.view, db.view, op, n, od, ob);
    count_launch();
    return check_launch("k_str_compare");
  });
}

int32_t tq_vec_string_unary(int32_t op, int64_t n, const tq_column *a, tq_column *out, int32_t mem) {
  TQ_TRY(ensure_init());
  if ((op != TQ_STR_LENGTH && op != TQ_STR_ISNULL) || n < 0) { set_error("bad string unary op %d", op); return TQ_ERR_INVALID_ARG; }
  TQ_TRY(check_str_col(a, n, "string argument"));
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  StrDev da;
  if (n) TQ_TRY(stage_str(a, n, mem, da, r.compute));
  return run_str(n, out, mem, [&](uint64_t *od, uint32_t *ob) -> int32_t {
    k_str_unary<<<str_grid(n), 256, 0, r.compute>>>(da.view, op, n, od, ob);
    count_launch();
    return check_launch("k_str_unary");
  });
}

int32_t tq_vec_in_string(int64_t n, const tq_column *a, int32_t n_list, const tq_column *list, tq_column *out, int32_t mem) {
  TQ_TRY(ensure_init());
  if (n < 0 || n_list < 0 || n_list > STR_MAX_IN_LIST || (n_list && !list)) { set_error("string IN list of %d columns (max %d per call)", n_list, STR_MAX_IN_LIST); return TQ_ERR_INVALID_ARG; }
  TQ_TRY(check_str_col(a, n, "string IN argument"));
  for (int j = 0; j < n_list; j++) TQ_TRY(check_str_col(&list[j], n, "string IN list element"));
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  StrDev da, dl[STR_MAX_IN_LIST];
  StrInList L{};
  L.n = n_list;
  if (n) {
    TQ_TRY(stage_str(a, n, mem, da, r.compute));
    for (int j = 0; j < n_list; j++) { TQ_TRY(stage_str(&list[j], n, mem, dl[j], r.compute)); L.c[j] = dl[j].view; }
  }
  return run_str(n, out, mem, [&](uint64_t *od, uint32_t *ob) -> int32_t {
    k_str_in<<<str_grid(((n + 31) >> 5) * 32), 256, 0, r.compute>>>(da.view, L, n, od, ob);
    count_launch();
    return check_launch("k_str_in");
  });
}

// mode 0: IF(cond, a, b)   mode 1: IFNULL(a, b).  out: offsets (n + 1) + data (capacity >= bytes(a) + bytes(b)) + null_bitmap.
static int32_t vec_pick_string(int mode, int64_t n, const tq_column *cond, const tq_column *a, const tq_column *b, tq_column *out, int32_t mem) {
  TQ_TRY(ensure_init());
  if (n < 0 || !out || (mode == 0 && (!cond || (n > 0 && !cond->data)))) return TQ_ERR_INVALID_ARG;
  TQ_TRY(check_str_col(a, n, "string IF / IFNULL argument"));
  TQ_TRY(check_str_col(b, n, "string IF / IFNULL argument"));
  if (n > 0 && (!out->offsets || !out->null_bitmap)) { set_error("string result column needs offsets and null_bitmap buffers"); return TQ_ERR_INVALID_ARG; }
  out->length = n;
  if (n == 0) { if (out->offsets) out->offsets[0] = 0; return TQ_OK; }
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  cudaStream_t s = r.compute;
  StrDev da, db;
  TQ_TRY(stage_str(a, n, mem, da, s));
  TQ_TRY(stage_str(b, n, mem, db, s));
  DevBuf dc, dcb, lens, off32, src, scan, tot, o_off, o_bm, o_bytes;
  const uint64_t *d_cond = nullptr;
  const uint32_t *d_cond_bm = nullptr;
  if (mode == 0) {
    if (mem == TQ_MEM_DEVICE) { d_cond = (const uint64_t *)cond->data; d_cond_bm = (const uint32_t *)cond->null_bitmap; }
    else {
      TQ_TRY(dc.reserve((size_t)n * 8));
      TQ_CUDA(cudaMemcpyAsync(dc.p, cond->data, (size_t)n * 8, cudaMemcpyHostToDevice, s));
      d_cond = dc.as<uint64_t>();
      if (cond->null_bitmap) {
        TQ_TRY(dcb.reserve(bitmap_alloc_bytes(n)));
        TQ_CUDA(cudaMemcpyAsync(dcb.p, cond->null_bitmap, bitmap_bytes(n), cudaMemcpyHostToDevice, s));
        d_cond_bm = dcb.as<uint32_t>();
      }
    }
  }
  TQ_TRY(lens.reserve((size_t)(n + 2) * 4));
  TQ_TRY(off32.reserve((size_t)(n + 2) * 4));
  TQ_TRY(src.reserve((size_t)n + 16));
  TQ_TRY(tot.reserve(16));
  uint32_t *bm_dev = (uint32_t *)out->null_bitmap;
  int64_t *off_dev = out->offsets;
  if (mem != TQ_MEM_DEVICE) {
    TQ_TRY(o_bm.reserve(bitmap_alloc_bytes(n)));
    TQ_TRY(o_off.reserve((size_t)(n + 1) * 8));
    bm_dev = o_bm.as<uint32_t>();
    off_dev = o_off.as<int64_t>();
  }
  k_str_pick<<<str_grid(n), 256, 0, s>>>(mode, d_cond, d_cond_bm, da.view, db.view, n, lens.as<uint32_t>(), src.as<uint8_t>(), bm_dev);
  count_launch();
  TQ_TRY(check_launch("k_str_pick"));
  TQ_TRY(exclusive_scan_u32(lens.as<uint32_t>(), 1, off32.as<uint32_t>(), 1, n, tot.as<uint64_t>(), scan, s));
  uint64_t total = 0;
  TQ_CUDA(cudaMemcpyAsync(&total, tot.p, 8, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaStreamSynchronize(s));
  if (total > 0xFFFFFFF0ull) { set_error("string result of one call exceeds 4 GiB"); return TQ_ERR_INVALID_ARG; }
  uint8_t *bytes_dev = out->data;
  if (mem != TQ_MEM_DEVICE) { TQ_TRY(o_bytes.reserve((size_t)total + 16)); bytes_dev = o_bytes.as<uint8_t>(); }
  if (total && !out->data) { set_error("string result column needs a data buffer of %llu bytes", (unsigned long long)total); return TQ_ERR_INVALID_ARG; }
  k_str_pick_copy<<<str_grid(n * 32 < (int64_t)1 << 24 ? n * 32 : (int64_t)1 << 24), 256, 0, s>>>(da.view, db.view, src.as<uint8_t>(), off32.as<uint32_t>(), tot.as<uint64_t>(), n,
                                                                                                off_dev, bytes_dev);
  count_launch();
  TQ_TRY(check_launch("k_str_pick_copy"));
  if (mem != TQ_MEM_DEVICE) {
    TQ_CUDA(cudaMemcpyAsync(out->offsets, off_dev, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
    TQ_CUDA(cudaMemcpyAsync(out->null_bitmap, bm_dev, bitmap_bytes(n), cudaMemcpyDeviceToHost, s));
    if (total) TQ_CUDA(cudaMemcpyAsync(out->data, bytes_dev, (size_t)total, cudaMemcpyDeviceToHost, s));
  }
  TQ_CUDA(cudaStreamSynchronize(s));
  return TQ_OK;
}

int32_t tq_vec_if_string(int64_t n, const tq_column *cond, const tq_column *a, const tq_column *b, tq_column *out, int32_t mem) {
  return vec_pick_string(0, n, cond, a, b, out, mem);
}
int32_t tq_vec_ifnull_string(int64_t n, const tq_column *a, const tq_column *b, tq_column *out, int32_t mem) {
  return vec_pick_string(1, n, nullptr, a, b, out, mem);
}

}  // extern "C"
