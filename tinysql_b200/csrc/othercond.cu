// othercond.cu — see othercond.cuh.  8-byte gathers / scatters over the result batch: HBM-bound.
#include "othercond.cuh"

namespace tq {

static int oc_grid(int64_t n) {
  const int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)rt().sm_count * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

// types/compare.go:44-100 VecCompare{II,UU,UI,IU}
__device__ __forceinline__ int oc_cmp_int(bool ua, bool ub, int64_t x, int64_t y) {
  if (ua && ub) { const uint64_t a = (uint64_t)x, b = (uint64_t)y; return a < b ? -1 : (a == b ? 0 : 1); }
  if (!ua && !ub) return x < y ? -1 : (x == y ? 0 : 1);
  if (ua) { if (y < 0 || (uint64_t)x > 0x7fffffffffffffffull) return 1; return x < y ? -1 : (x == y ? 0 : 1); }
  if (x < 0 || (uint64_t)y > 0x7fffffffffffffffull) return -1;
  return x < y ? -1 : (x == y ? 0 : 1);
}

__device__ __forceinline__ bool oc_row_passes(const OcPlan &pl, const OcCols &c, int64_t i) {
  for (int k = 0; k < pl.n_conds; k++) {
    const OcCond &q = pl.c[k];
    if (!tqd::bm_not_null(c.bm[q.lhs], i)) return false;
    const uint64_t a = c.data[q.lhs][i];
    uint64_t b = q.cbits;
    if (q.rhs >= 0) {
      if (!tqd::bm_not_null(c.bm[q.rhs], i)) return false;
      b = c.data[q.rhs][i];
    }
    int cmp;
    if (q.lhs_type == TQ_TYPE_FLOAT64) {
      const double x = __longlong_as_double((long long)a), y = __longlong_as_double((long long)b);
      cmp = x < y ? -1 : (x == y ? 0 : 1);
    } else {
      cmp = oc_cmp_int(q.lhs_type == TQ_TYPE_UINT64, q.rhs_type == TQ_TYPE_UINT64, (int64_t)a, (int64_t)b);
    }
    bool ok;
    switch (q.op) {
      case TQ_CMP_LT: ok = cmp < 0; break;
      case TQ_CMP_LE: ok = cmp <= 0; break;
      case TQ_CMP_GT: ok = cmp > 0; break;
      case TQ_CMP_GE: ok = cmp >= 0; break;
      case TQ_CMP_EQ: ok = cmp == 0; break;
      default: ok = cmp != 0; break;
    }
    if (!ok) return false;
  }
  return true;
}

// flag: 0 = miss row of the outer join (kept as is), 1 = key match that passes, 2 = key match that fails
__global__ void __launch_bounds__(256) k_oc_eval(const OcPlan pl, const OcCols c, int64_t n, uint8_t *flag, uint32_t *surv, uint32_t *first) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool match = pl.outer ? tqd::bm_not_null(c.bm[pl.build_key_col], i) : true;
    uint8_t f = 0;
    if (match) {
      const bool pass = oc_row_passes(pl, c, i);
      f = pass ? 1 : 2;
      if (pl.outer) {
        const uint32_t pid = (uint32_t)c.data[pl.rowid_col][i];
        if (pass) atomicAdd(&surv[pid], 1u);
        else atomicMin(&first[pid], (uint32_t)i);
      }
    }
    flag[i] = f;
  }
}

// keep[i] = 1 for rows that stay; flag 3 marks the ONE failed row of a survivor-less probe row that becomes its miss row
__global__ void __launch_bounds__(256) k_oc_decide(const OcPlan pl, const OcCols c, int64_t n, uint8_t *flag, const uint32_t *surv, const uint32_t *first,
                                                    uint32_t *keep) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint8_t f = flag[i];
    uint32_t k = (f == 0 || f == 1) ? 1u : 0u;
    if (f == 2 && pl.outer) {
      const uint32_t pid = (uint32_t)c.data[pl.rowid_col][i];
      if (surv[pid] == 0 && first[pid] == (uint32_t)i) { f = 3; k = 1; flag[i] = 3; }
    }
    keep[i] = k;
  }
}

__global__ void __launch_bounds__(256) k_oc_compact(const OcPlan pl, const OcCols c, int64_t n, const uint8_t *flag, const uint32_t *keep, const uint32_t *pos) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (!keep[i]) continue;
    const uint32_t d = pos[i];
    const bool to_miss = flag[i] == 3;
    for (int col = 0; col < c.n; col++) {
      const bool build_side = col >= pl.build_lo && col < pl.build_hi;
      if (to_miss && build_side) {   // onMissMatch: the inner side becomes defaultInner
        const int bc = col - pl.build_lo;
        const bool dn = (pl.def_mask >> bc) & 1u;
        c.out_data[col][d] = dn ? pl.def_val[bc] : 0ull;
        if (dn) atomicOr(&c.out_bm[col][d >> 5], 1u << (d & 31));
        continue;
      }
      const bool nn = tqd::bm_not_null(c.bm[col], i);
      c.out_data[col][d] = nn ? c.data[col][i] : 0ull;
      if (nn) atomicOr(&c.out_bm[col][d >> 5], 1u << (d & 31));
    }
  }
}

int32_t oc_filter(const OcPlan &plan, const OcCols &cols, int64_t n, int64_t n_probe_rows, DevBuf &scratch, DevBuf &scan_scratch, int64_t *n_out,
                  cudaStream_t s) {
  *n_out = 0;
  if (n <= 0) return TQ_OK;
  if (n > 0xFFFFFFF0ll || n_probe_rows > 0xFFFFFFF0ll) { set_error("OtherConditions: result batch too large"); return TQ_ERR_INVALID_ARG; }
  // scratch: keep u32[n] | pos u32[n] | surv u32[np] | first u32[np] | total u64 | flag u8[n]
  const size_t np = plan.outer ? (size_t)n_probe_rows : 0;
  const size_t words = (size_t)n * 2 + np * 2;
  const size_t total_off = (words * 4 + 7) & ~(size_t)7;
  TQ_TRY(scratch.reserve(total_off + 8 + (size_t)n + 16));
  uint32_t *keep = scratch.as<uint32_t>(), *pos = keep + n, *surv = pos + n, *first = surv + np;
  uint64_t *d_total = reinterpret_cast<uint64_t *>(scratch.as<uint8_t>() + total_off);
  uint8_t *flag = scratch.as<uint8_t>() + total_off + 8;
  if (np) {
    TQ_CUDA(cudaMemsetAsync(surv, 0, np * 4, s));
    TQ_CUDA(cudaMemsetAsync(first, 0xFF, np * 4, s));
  }
  k_oc_eval<<<oc_grid(n), 256, 0, s>>>(plan, cols, n, flag, surv, first);
  k_oc_decide<<<oc_grid(n), 256, 0, s>>>(plan, cols, n, flag, surv, first, keep);
  count_launch(2);
  TQ_TRY(check_launch("k_oc_decide"));
  TQ_TRY(exclusive_scan_u32(keep, 1, pos, 1, n, d_total, scan_scratch, s));
  for (int c = 0; c < cols.n; c++) TQ_CUDA(cudaMemsetAsync(cols.out_bm[c], 0, bitmap_alloc_bytes(n), s));
  k_oc_compact<<<oc_grid(n), 256, 0, s>>>(plan, cols, n, flag, keep, pos);
  count_launch();
  TQ_TRY(check_launch("k_oc_compact"));
  uint64_t total = 0;
  TQ_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaStreamSynchronize(s));
  *n_out = (int64_t)total;
  return TQ_OK;
}

}  // namespace tq
