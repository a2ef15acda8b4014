// varlen.cu — row-id gather of FLOAT / var-len join payload cells (see varlen.cuh).  Pure byte movement: HBM-bound.
#include "varlen.cuh"

namespace tq {

static int vl_grid(int64_t n) {
  const int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)rt().sm_count * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

__global__ void k_iota_u64(uint64_t *dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (uint64_t)i;
}

__global__ void __launch_bounds__(256) k_gather_f32(const uint32_t *src, const uint64_t *rowids, const uint32_t *bm, int64_t n, uint32_t *out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = tqd::bm_not_null(bm, i) ? src[rowids[i]] : 0u;
}

__global__ void __launch_bounds__(256) k_var_lens(const int64_t *off, const uint64_t *rowids, const uint32_t *bm, int64_t n, uint32_t *lens) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t len = 0;
    if (tqd::bm_not_null(bm, i)) {
      const uint64_t id = rowids[i];
      len = (uint32_t)(off[id + 1] - off[id]);
    }
    lens[i] = len;
  }
}

// one warp per output cell; 16-byte body when source and destination are mutually aligned
__global__ void __launch_bounds__(256) k_var_copy(const int64_t *off, int64_t base, const uint8_t *src_bytes, const uint64_t *rowids, const uint32_t *bm,
                                                   int64_t n, const uint32_t *out_off32, uint64_t total, int64_t *out_off, uint8_t *out_bytes) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (warp == 0 && lane == 0) out_off[n] = (int64_t)total;
  for (int64_t i = warp; i < n; i += n_warps) {
    const uint32_t o = out_off32[i];
    if (lane == 0) out_off[i] = (int64_t)o;
    if (!tqd::bm_not_null(bm, i)) continue;
    const uint64_t id = rowids[i];
    const int64_t s0 = off[id], s1 = off[id + 1];
    const uint8_t *src = src_bytes + (s0 - base);
    uint8_t *dst = out_bytes + o;
    const int64_t len = s1 - s0;
    int64_t done = 0;
    if ((((uintptr_t)src ^ (uintptr_t)dst) & 15) == 0 && len >= 64) {
      int64_t head = (int64_t)((16 - ((uintptr_t)dst & 15)) & 15);
      for (int64_t b = lane; b < head; b += 32) dst[b] = src[b];
      const int64_t vecs = (len - head) >> 4;
      const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
      uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
      for (int64_t v = lane; v < vecs; v += 32) d4[v] = s4[v];
      done = head + (vecs << 4);
    }
    for (int64_t b = done + lane; b < len; b += 32) dst[b] = src[b];
  }
}

__global__ void k_widen_f32(const uint32_t *src, int64_t n, uint64_t *dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = (uint64_t)__double_as_longlong((double)__uint_as_float(src[i]));
}
__global__ void k_narrow_f64(const uint64_t *src, int64_t n, uint32_t *dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float_as_uint((float)__longlong_as_double((long long)src[i]));
}

int32_t widen_f32(const uint32_t *src, int64_t n, uint64_t *dst, cudaStream_t s) {
  if (n <= 0) return TQ_OK;
  k_widen_f32<<<vl_grid(n), 256, 0, s>>>(src, n, dst);
  count_launch();
  return check_launch("k_widen_f32");
}
int32_t narrow_f64(const uint64_t *src, int64_t n, uint32_t *dst, cudaStream_t s) {
  if (n <= 0) return TQ_OK;
  k_narrow_f64<<<vl_grid(n), 256, 0, s>>>(src, n, dst);
  count_launch();
  return check_launch("k_narrow_f64");
}

// Upload the staged cells of an indirect column into its device store.
int32_t upload_store(const HostVarAccum &h, SideStore &st, cudaStream_t s) {
  st.elem = h.elem;
  st.n = h.n;
  st.base = 0;
  st.nbytes = (int64_t)h.bytes.size();
  TQ_TRY(st.bytes.reserve(h.bytes.size() + 16));
  if (!h.bytes.empty()) TQ_CUDA(cudaMemcpyAsync(st.bytes.p, h.bytes.data(), h.bytes.size(), cudaMemcpyHostToDevice, s));
  if (h.elem == 0) {
    TQ_TRY(st.offsets.reserve(h.off.size() * 8));
    TQ_CUDA(cudaMemcpyAsync(st.offsets.p, h.off.data(), h.off.size() * 8, cudaMemcpyHostToDevice, s));
  }
  TQ_CUDA(cudaStreamSynchronize(s));  // pageable source
  return TQ_OK;
}

int32_t iota_u64(uint64_t *dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return TQ_OK;
  k_iota_u64<<<vl_grid(n), 256, 0, s>>>(dst, n);
  count_launch();
  return check_launch("k_iota_u64");
}

int32_t gather_cells(const SideStore &st, const uint64_t *rowids, const uint32_t *bm, int64_t n, VarOut &out, DevBuf &lens, DevBuf &scan_scratch,
                     cudaStream_t s) {
  out.used = true;
  out.on_host = false;
  out.elem = st.elem;
  out.total = 0;
  if (st.elem == 4) {
    TQ_TRY(out.bytes.reserve((size_t)(n ? n : 1) * 4));
    if (n > 0) {
      k_gather_f32<<<vl_grid(n), 256, 0, s>>>(st.bytes.as<uint32_t>(), rowids, bm, n, out.bytes.as<uint32_t>());
      count_launch();
      TQ_TRY(check_launch("k_gather_f32"));
    }
    out.total = n * 4;
    TQ_CUDA(cudaStreamSynchronize(s));
    return TQ_OK;
  }
  TQ_TRY(out.off.reserve((size_t)(n + 1) * 8));
  if (n == 0) {
    TQ_CUDA(cudaMemsetAsync(out.off.p, 0, 8, s));
    TQ_TRY(out.bytes.reserve(16));
    TQ_CUDA(cudaStreamSynchronize(s));
    return TQ_OK;
  }
  // lens | exclusive offsets (u32) | total (u64) share one scratch buffer
  const size_t words = (size_t)((n + 1 + 1) & ~1ull);
  TQ_TRY(lens.reserve(words * 8 + 16));
  uint32_t *d_len = lens.as<uint32_t>(), *d_off32 = d_len + words;
  uint64_t *d_total = reinterpret_cast<uint64_t *>(d_off32 + words);
  k_var_lens<<<vl_grid(n), 256, 0, s>>>(st.offsets.as<int64_t>(), rowids, bm, n, d_len);
  count_launch();
  TQ_TRY(check_launch("k_var_lens"));
  TQ_TRY(exclusive_scan_u32(d_len, 1, d_off32, 1, n, d_total, scan_scratch, s));
  uint64_t total = 0;
  TQ_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaStreamSynchronize(s));
  if (total > 0xFFFFFFF0ull) { set_error("var-len output of one result batch exceeds 4 GiB (%llu bytes): use a smaller probe_batch_rows", (unsigned long long)total); return TQ_ERR_INVALID_ARG; }
  TQ_TRY(out.bytes.reserve((size_t)total + 16));
  const int64_t warps_needed = n;
  int64_t blocks = (warps_needed * 32 + 255) / 256;
  const int64_t cap = (int64_t)rt().sm_count * 16;
  if (blocks > cap) blocks = cap;
  k_var_copy<<<(int)blocks, 256, 0, s>>>(st.offsets.as<int64_t>(), st.base, st.bytes.as<uint8_t>(), rowids, bm, n, d_off32, total, out.off.as<int64_t>(),
                                        out.bytes.as<uint8_t>());
  count_launch();
  TQ_TRY(check_launch("k_var_copy"));
  out.total = (int64_t)total;
  TQ_CUDA(cudaStreamSynchronize(s));
  return TQ_OK;
}

}  // namespace tq
