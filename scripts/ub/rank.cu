// microbenchmark: cost of ranking 7-bit bin ids inside a warp / CTA — ballots vs match.any vs shared-memory atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix64(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
template <int MODE>
__global__ void __launch_bounds__(512) k(const uint64_t *keys, int64_t n, int nbits, unsigned *out) {
  __shared__ unsigned s_cnt[16 * 160];
  for (int i = threadIdx.x; i < 16 * 160; i += 512) s_cnt[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  unsigned acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * 512 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 512) {
    const uint64_t key = keys[i];
    unsigned b;
    if (MODE == 4) b = (unsigned)(((key ^ (key >> 32)) * 0x9E3779B97F4A7C15ull) >> (64 - nbits));
    else b = (unsigned)(mix64(key) >> (64 - nbits));
    unsigned rank;
    if (MODE == 0 || MODE == 4) {  // ballots
      unsigned peers = 0xffffffffu;
      for (int bit = 0; bit < nbits; bit++) { const bool one = (b >> bit) & 1u; const unsigned v = __ballot_sync(0xffffffffu, one); peers &= one ? v : ~v; }
      const int leader = __ffs(peers) - 1;
      unsigned old = 0;
      if (lane == leader) { old = s_cnt[warp * 160 + b]; s_cnt[warp * 160 + b] = old + __popc(peers); }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank = old + __popc(peers & lt);
      __syncwarp();
    } else if (MODE == 1) {  // match.any
      const unsigned peers = __match_any_sync(0xffffffffu, b);
      const int leader = __ffs(peers) - 1;
      unsigned old = 0;
      if (lane == leader) { old = s_cnt[warp * 160 + b]; s_cnt[warp * 160 + b] = old + __popc(peers); }
      old = __shfl_sync(0xffffffffu, old, leader);
      rank = old + __popc(peers & lt);
      __syncwarp();
    } else if (MODE == 2) {  // shared atomic with return, one histogram per CTA
      rank = atomicAdd(&s_cnt[b], 1u);
    } else {  // MODE 3: shared atomic with return, one histogram per warp
      rank = atomicAdd(&s_cnt[warp * 160 + b], 1u);
    }
    acc += rank;
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}
int main() {
  const int64_t n = 1ll << 27;
  uint64_t *keys; unsigned *out;
  cudaMalloc(&keys, n * 8); cudaMalloc(&out, 4);
  uint64_t *h = (uint64_t *)malloc(n * 8);
  uint64_t x = 88172645463325252ull;
  for (int64_t i = 0; i < n; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = x; }
  cudaMemcpy(keys, h, n * 8, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const char *names[5] = {"ballot x nbits + mix64", "match.any + mix64", "ATOMS ret (CTA hist) + mix64", "ATOMS ret (warp hist) + mix64", "ballot x nbits + mul hash"};
  for (int nbits = 5; nbits <= 8; nbits += 3)
    for (int mode = 0; mode < 5; mode++)
      for (int ctas = 1; ctas <= 4; ctas *= 2) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
          cudaEventRecord(a);
          if (mode == 0) k<0><<<148 * ctas, 512>>>(keys, n, nbits, out);
          if (mode == 1) k<1><<<148 * ctas, 512>>>(keys, n, nbits, out);
          if (mode == 2) k<2><<<148 * ctas, 512>>>(keys, n, nbits, out);
          if (mode == 3) k<3><<<148 * ctas, 512>>>(keys, n, nbits, out);
          if (mode == 4) k<4><<<148 * ctas, 512>>>(keys, n, nbits, out);
          cudaEventRecord(b); cudaEventSynchronize(b);
          float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("nbits=%d %-32s ctas/SM=%d  %.3f ms for %lld rows  (%.2f ms per 1e8 rows; key read alone = %.2f ms at 6.5 TB/s)\n", nbits, names[mode], ctas, best, (long long)n, best * 1e8 / n, 0.8 / 6.5);
      }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
