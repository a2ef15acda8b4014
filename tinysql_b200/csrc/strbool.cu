// strbool.cu — toBool of an ETString filter expression (expression/expression.go:308-322), the last evaluation type of
// VectorizedFilter / VecEvalBool on the device: selected[i] = cell i is not NULL and types.StrToInt(cell) != 0.
// One thread per row walks its cell once (strnum.cuh).  The error VecEvalBool reports is the error of the LAST non-NULL row
// (`err = err1` inside the loop overwrites earlier ones): the kernel keeps max(row << 1 | failed) and the host looks at bit 0.
#include "common.cuh"
#include "strnum.cuh"

using namespace tq;

namespace {

__global__ void __launch_bounds__(256) k_filter_string(const int64_t *__restrict__ off, const uint8_t *__restrict__ data, int64_t base, const uint32_t *__restrict__ bm, int64_t n,
                                                       uint8_t *__restrict__ selected, unsigned long long *last) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint8_t sel = 0;
    if (tqd::bm_not_null(bm, i)) {
      int err = 0;
      const int64_t v = tqd::str_to_int(data + (off[i] - base), off[i + 1] - off[i], &err);
      sel = v != 0;
      mine = ((unsigned long long)(i + 1) << 1) | (unsigned long long)(err != 0);   // rows ascend per thread: the last assignment is this thread's last row
    }
    selected[i] = sel;
  }
  if (mine) atomicMax(last, mine);
}

}  // namespace

extern "C" int32_t tq_vec_filter_string(int64_t n, const tq_column *a, uint8_t *selected, int32_t mem) {
  TQ_TRY(ensure_init());
  if (n < 0 || !a || (n > 0 && (!a->offsets || !selected))) { set_error("string filter: a var-len column (offsets + data) and a selected buffer are required"); return TQ_ERR_INVALID_ARG; }
  if (n == 0) return TQ_OK;
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  cudaStream_t s = r.compute;
  const int64_t blocks = (n + 255) / 256;
  const int grid = (int)(blocks < (int64_t)r.sm_count * 8 ? blocks : (int64_t)r.sm_count * 8);
  DevBuf d_off, d_data, d_bm, d_sel, d_last;
  TQ_TRY(d_last.reserve(8));
  TQ_CUDA(cudaMemsetAsync(d_last.p, 0, 8, s));
  const int64_t *off = a->offsets;
  const uint8_t *data = a->data;
  const uint32_t *bm = (const uint32_t *)a->null_bitmap;
  int64_t base = 0;
  uint8_t *sel = selected;
  if (mem != TQ_MEM_DEVICE) {
    const int64_t b0 = a->offsets[0], b1 = a->offsets[n];
    if (b1 < b0 || (b1 > b0 && !a->data)) { set_error("malformed var-len column"); return TQ_ERR_INVALID_ARG; }
    TQ_TRY(d_off.reserve((size_t)(n + 1) * 8));
    TQ_TRY(d_data.reserve((size_t)(b1 - b0) + 16));
    TQ_TRY(d_sel.reserve((size_t)n));
    TQ_CUDA(cudaMemcpyAsync(d_off.p, a->offsets, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s));
    if (b1 > b0) TQ_CUDA(cudaMemcpyAsync(d_data.p, a->data + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, s));
    bm = nullptr;
    if (a->null_bitmap) {
      TQ_TRY(d_bm.reserve(bitmap_alloc_bytes(n)));
      TQ_CUDA(cudaMemcpyAsync(d_bm.p, a->null_bitmap, bitmap_bytes(n), cudaMemcpyHostToDevice, s));
      bm = d_bm.as<uint32_t>();
    }
    off = d_off.as<int64_t>();
    data = d_data.as<uint8_t>();
    base = b0;
    sel = d_sel.as<uint8_t>();
  }
  TQ_LAUNCH(k_filter_string, grid, 256, 0, s, off, data, base, bm, n, sel, d_last.as<unsigned long long>());
  count_launch();
  TQ_TRY(check_launch("k_filter_string"));
  unsigned long long last = 0;
  if (mem != TQ_MEM_DEVICE) TQ_CUDA(cudaMemcpyAsync(selected, d_sel.p, (size_t)n, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaMemcpyAsync(&last, d_last.p, 8, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaStreamSynchronize(s));
  if (last & 1ull) {   // types.ErrOverflow.GenWithStackByArgs("BIGINT", validPrefix) of the last non-NULL row
    set_error("BIGINT value is out of range in a string filter expression (row %llu)", (last >> 1) - 1);
    return TQ_ERR_OVERFLOW_BIGINT;
  }
  return TQ_OK;
}
