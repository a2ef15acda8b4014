// join.cu — HashJoinExec on the device (replaces executor/join.go, hash_table.go, joiner.go).
//
// Build (tq_join_finalize_build): the inner side is materialised in HBM, then
//   k_build_insert   every non-NULL-key row finds/claims the slot of its key in an open-addressed
//                    table of DISTINCT keys (atomicCAS) and bumps the slot's count
//   exclusive scan   slot counts -> CSR offsets
//   k_build_fill     row ids are scattered into their key's CSR segment
//   k_build_fixsort  offsets restored, segments of duplicate keys sorted ascending (= the reference's
//                    insertion order, rowHashMap.Get hash_table.go:259-272)
//   k_gather_col     build columns are permuted into CSR order, so a probe hit (off, cnt) addresses
//                    cnt CONTIGUOUS build rows — no row-pointer chasing on the probe side
// Probe (one launch per device batch): k_probe — each CTA takes 1024-row tiles: 128-bit slot loads,
//   per-row match counts, block scan, one atomicAdd for the tile's output range, then an
//   output-centric expansion (thread per OUTPUT row, binary search in the tile's prefix array) with
//   coalesced 8-byte column stores and ballot-assembled null-bitmap words.
#include <deque>
#include <memory>
#include <new>

#include "common.cuh"

namespace tq {

static constexpr int MAXC = 16;  // columns per join side
static constexpr uint64_t EMPTY_KEY = 0xA5C3F00DDEADBEEFull;  // slot sentinel; a real key with this value lives in a side segment
static constexpr uint32_t ROW_INVALID = 0xFFFFFFFFu, ROW_SENTINEL = 0xFFFFFFFEu;
static constexpr uint32_t OFF_MISS = 0xFFFFFFFFu;

struct __align__(16) Slot {
  uint64_t key;
  uint32_t off;
  uint32_t cnt;
};

// key_mode: how (flag, raw bytes) equality (util/codec/codec.go:212-240,363-382) maps onto raw 8-byte equality
//   0: flags always agree (both signed, both unsigned, or both DOUBLE)  -> raw equality
//   1: one side UNSIGNED, the other signed: values with the sign bit set carry different flags
//      (uvarintFlag vs varintFlag) and can never match -> such rows are treated like NULL keys
//   2: int-class vs DOUBLE: flags never agree -> nothing matches
enum { KEYMODE_RAW = 0, KEYMODE_NO_SIGNBIT = 1, KEYMODE_NEVER = 2 };

struct JoinTable {
  Slot *slots;
  uint64_t mask;
  uint32_t sent_off, sent_cnt;
};

__device__ __forceinline__ bool key_valid(uint64_t key, bool not_null, int key_mode) {
  if (!not_null) return false;
  if (key_mode == KEYMODE_RAW) return true;
  if (key_mode == KEYMODE_NO_SIGNBIT) return (key >> 63) == 0;
  return false;
}

__global__ void k_init_slots(Slot *slots, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Slot s;
    s.key = EMPTY_KEY; s.off = 0; s.cnt = 0;
    slots[i] = s;
  }
}

// counters[0] = sentinel-key row count, [1] = sentinel fill cursor, [2] = distinct keys, [3] = large-segment worklist length
__global__ void __launch_bounds__(256) k_build_insert(const uint64_t *keys, const uint32_t *bm, int64_t n, int key_mode, Slot *slots,
                                                       uint64_t mask, uint32_t *row_slot, uint32_t *counters) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint64_t key = keys[i];
    if (!key_valid(key, tqd::bm_not_null(bm, i), key_mode)) { row_slot[i] = ROW_INVALID; continue; }  // hash_table.go:161-163
    if (key == EMPTY_KEY) { atomicAdd(&counters[0], 1u); row_slot[i] = ROW_SENTINEL; continue; }
    uint64_t idx = tqd::mix64(key) & mask;
    for (;;) {
      const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(&slots[idx].key), (unsigned long long)EMPTY_KEY,
                                                (unsigned long long)key);
      if (prev == EMPTY_KEY || prev == key) {
        atomicAdd(&slots[idx].cnt, 1u);
        row_slot[i] = (uint32_t)idx;
        break;
      }
      idx = (idx + 1) & mask;
    }
  }
}

__global__ void __launch_bounds__(256) k_build_fill(const uint32_t *row_slot, int64_t n, Slot *slots, uint32_t sent_off, uint32_t *counters,
                                                     uint32_t *row_ids) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint32_t s = row_slot[i];
    if (s == ROW_INVALID) continue;
    uint32_t pos;
    if (s == ROW_SENTINEL) pos = sent_off + atomicAdd(&counters[1], 1u);
    else pos = atomicAdd(&slots[s].off, 1u);
    row_ids[pos] = (uint32_t)i;
  }
}

// Restores off (k_build_fill advanced it by cnt) and sorts duplicate-key segments ascending by row id.
__global__ void __launch_bounds__(256) k_build_fixsort(Slot *slots, uint64_t n_slots, uint32_t *row_ids, uint32_t *counters,
                                                        uint2 *worklist, uint32_t worklist_cap) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned distinct = 0;
  for (; i < n_slots; i += stride) {
    const uint32_t cnt = slots[i].cnt;
    if (cnt == 0) continue;
    distinct++;
    const uint32_t off = slots[i].off - cnt;
    slots[i].off = off;
    if (cnt == 1) continue;
    if (cnt <= 32) {
      uint32_t *seg = row_ids + off;  // insertion sort: segments are tiny
      for (uint32_t a = 1; a < cnt; a++) {
        const uint32_t v = seg[a];
        uint32_t b = a;
        while (b > 0 && seg[b - 1] > v) { seg[b] = seg[b - 1]; b--; }
        seg[b] = v;
      }
    } else {
      const uint32_t w = atomicAdd(&counters[3], 1u);
      if (w < worklist_cap) worklist[w] = make_uint2(off, cnt);
    }
  }
  distinct = __reduce_add_sync(0xffffffffu, distinct);
  if ((threadIdx.x & 31) == 0 && distinct) atomicAdd(&counters[2], distinct);
}

// One CTA per large duplicate segment: bitonic sort in global memory (indices >= cnt act as +inf).
__global__ void __launch_bounds__(256) k_sort_large(const uint2 *worklist, uint32_t *row_ids) {
  const uint2 w = worklist[blockIdx.x];
  uint32_t *seg = row_ids + w.x;
  const uint32_t n = w.y;
  uint32_t p2 = 1;
  while (p2 < n) p2 <<= 1;
  // all comparators ascending (first stage of each merge mirrors: partner = t ^ (k-1)), so the virtual
  // +inf tail never has to move and comparators touching it are skipped
  for (uint32_t k = 2; k <= p2; k <<= 1) {
    for (uint32_t t = threadIdx.x; t < p2; t += blockDim.x) {
      const uint32_t partner = t ^ (k - 1);
      if (partner > t && partner < n) {
        const uint32_t a = seg[t], b = seg[partner];
        if (a > b) { seg[t] = b; seg[partner] = a; }
      }
    }
    __syncthreads();
    for (uint32_t j = k >> 2; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < p2; t += blockDim.x) {
        const uint32_t partner = t ^ j;
        if (partner > t && partner < n) {
          const uint32_t a = seg[t], b = seg[partner];
          if (a > b) { seg[t] = b; seg[partner] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// B'[pos] = B[row_ids[pos]] (data + null bit): build columns in CSR order.
__global__ void __launch_bounds__(256) k_gather_col(const uint64_t *src, const uint32_t *src_bm, const uint32_t *row_ids, int64_t n,
                                                     uint64_t *dst, uint32_t *dst_bm) {
  const int lane = threadIdx.x & 31;
  const int64_t n_words = (n + 31) >> 5;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_words; w += stride) {
    const int64_t pos = w * 32 + lane;
    bool nn = false;
    if (pos < n) {
      const uint32_t r = row_ids[pos];
      dst[pos] = src[r];
      nn = tqd::bm_not_null(src_bm, r);
    }
    const unsigned word = __ballot_sync(0xffffffffu, nn);
    if (dst_bm && lane == 0) dst_bm[w] = word;
  }
}

// ------------------------------------------------------------------ probe
static constexpr int PROBE_THREADS = 256;
static constexpr int PROBE_ROWS_PER_THREAD = 4;
static constexpr int PROBE_TILE = PROBE_THREADS * PROBE_ROWS_PER_THREAD;

struct ProbeParams {
  int n_probe_cols, n_build_cols;
  DCol probe[MAXC];
  DCol build[MAXC];           // CSR-ordered build columns
  DColMut out_probe[MAXC];    // destination of probe column c (bm == nullptr: column cannot hold NULLs, bitmap pre-filled)
  DColMut out_build[MAXC];
  const uint8_t *selected;    // outerSideFilter result or nullptr
  int key_col;
  int key_mode;
  int is_outer;               // LeftOuter / RightOuter: misses emit probe row ++ NULLs (joiner.go:274-277,337-340)
  int64_t n;
  uint64_t capacity;          // rows the output columns can hold
  unsigned long long *cursor; // [0] rows produced (may exceed capacity: then the batch is re-run), [1] matched probe rows
};

__device__ __forceinline__ Slot ld_slot(const Slot *p) {
  const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);  // one 128-bit load: key | off,cnt
  Slot s;
  s.key = v.x;
  s.off = (uint32_t)v.y;
  s.cnt = (uint32_t)(v.y >> 32);
  return s;
}

__global__ void __launch_bounds__(PROBE_THREADS) k_probe(const ProbeParams p, const JoinTable t) {
  __shared__ unsigned long long s_prefix[PROBE_TILE + 1];
  __shared__ uint32_t s_off[PROBE_TILE];
  __shared__ unsigned long long s_warp_sums[PROBE_THREADS / 32 + 1];
  __shared__ unsigned long long s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_tiles = (p.n + PROBE_TILE - 1) / PROBE_TILE;
  const uint64_t *keys = p.probe[p.key_col].data;
  const uint32_t *kbm = p.probe[p.key_col].bm;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t tile_base = tile * PROBE_TILE;
    // ---- phase A: look up PROBE_ROWS_PER_THREAD keys per thread (coalesced: row = base + k*256 + tid)
    uint64_t key[PROBE_ROWS_PER_THREAD];
    bool valid[PROBE_ROWS_PER_THREAD];
    Slot s[PROBE_ROWS_PER_THREAD];
    uint64_t idx[PROBE_ROWS_PER_THREAD];
#pragma unroll
    for (int k = 0; k < PROBE_ROWS_PER_THREAD; k++) {
      const int64_t r = tile_base + k * PROBE_THREADS + tid;
      key[k] = 0;
      valid[k] = false;
      if (r < p.n) {
        key[k] = tqd::ld_stream_u64(keys + r);
        const bool sel = p.selected ? (p.selected[r] != 0) : true;          // join.go:344 `!selected[i] || hasNull[i]` -> miss
        valid[k] = sel && key_valid(key[k], tqd::bm_not_null(kbm, r), p.key_mode);
      }
      idx[k] = tqd::mix64(key[k]) & t.mask;
    }
#pragma unroll
    for (int k = 0; k < PROBE_ROWS_PER_THREAD; k++) {  // the 4 random 16-byte loads are issued back to back
      if (valid[k] && key[k] != EMPTY_KEY) s[k] = ld_slot(t.slots + idx[k]);
      else { s[k].key = EMPTY_KEY; s[k].off = 0; s[k].cnt = 0; }
    }
#pragma unroll
    for (int k = 0; k < PROBE_ROWS_PER_THREAD; k++) {
      uint32_t off = OFF_MISS, cnt = 0;
      if (valid[k]) {
        if (key[k] == EMPTY_KEY) {
          if (t.sent_cnt) { off = t.sent_off; cnt = t.sent_cnt; }
        } else {
          Slot cur = s[k];
          uint64_t i = idx[k];
          while (cur.key != key[k] && cur.key != EMPTY_KEY) {  // linear probing; rare at load factor <= 0.5
            i = (i + 1) & t.mask;
            cur = ld_slot(t.slots + i);
          }
          if (cur.key == key[k]) { off = cur.off; cnt = cur.cnt; }
        }
      }
      const int64_t r = tile_base + k * PROBE_THREADS + tid;
      const uint32_t c = cnt ? cnt : ((p.is_outer && r < p.n) ? 1u : 0u);  // onMissMatch
      s_off[k * PROBE_THREADS + tid] = cnt ? off : OFF_MISS;
      s_prefix[k * PROBE_THREADS + tid] = c;
    }
    __syncthreads();
    // ---- phase B: exclusive scan of the tile's 1024 counts (thread owns 4 consecutive entries)
    unsigned long long c4[PROBE_ROWS_PER_THREAD], tsum = 0;
    unsigned matched = 0;
#pragma unroll
    for (int k = 0; k < PROBE_ROWS_PER_THREAD; k++) {
      c4[k] = s_prefix[tid * PROBE_ROWS_PER_THREAD + k];
      tsum += c4[k];
      matched += (s_off[tid * PROBE_ROWS_PER_THREAD + k] != OFF_MISS);
    }
    unsigned long long inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    matched = __reduce_add_sync(0xffffffffu, matched);
    if (lane == 31) s_warp_sums[warp] = inc;
    __syncthreads();  // also orders the c4 reads before the prefix writes below
    if (warp == 0) {
      unsigned long long w = (lane < PROBE_THREADS / 32) ? s_warp_sums[lane] : 0;
      unsigned long long winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += v;
      }
      if (lane < PROBE_THREADS / 32) s_warp_sums[lane] = winc - w;
      if (lane == PROBE_THREADS / 32 - 1) s_warp_sums[PROBE_THREADS / 32] = winc;
    }
    if (lane == 0 && matched) atomicAdd(p.cursor + 1, (unsigned long long)matched);
    __syncthreads();
    unsigned long long run = inc - tsum + s_warp_sums[warp];
#pragma unroll
    for (int k = 0; k < PROBE_ROWS_PER_THREAD; k++) {
      s_prefix[tid * PROBE_ROWS_PER_THREAD + k] = run;
      run += c4[k];
    }
    const unsigned long long M = s_warp_sums[PROBE_THREADS / 32];
    if (tid == 0) {
      s_prefix[PROBE_TILE] = M;
      s_base = M ? atomicAdd(p.cursor, M) : 0ull;  // ---- phase C: claim [base, base+M) of the output
    }
    __syncthreads();
    const unsigned long long base = s_base;
    // ---- phase D: output-centric expansion (skipped if this tile does not fit: the host re-runs the batch)
    if (M && base + M <= p.capacity) {
      const unsigned long long base_al = base & ~31ull;
      const unsigned long long end = base + M;
      for (unsigned long long q0 = base_al + (unsigned long long)warp * 32; q0 < end; q0 += PROBE_THREADS) {
        const unsigned long long q = q0 + lane;
        const bool active = q >= base && q < end;
        const bool full_word = q0 >= base && q0 + 32 <= end;
        int r = 0;
        unsigned long long j = 0;
        uint32_t off = OFF_MISS;
        if (active) {
          const unsigned long long o = q - base;
          int lo = 0, hi = PROBE_TILE;  // find r: prefix[r] <= o < prefix[r+1]
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_prefix[mid] <= o) lo = mid; else hi = mid;
          }
          r = lo;
          j = o - s_prefix[r];
          off = s_off[r];
        }
        // physical row of entry r: entries are stored [k*256 + tid]
        const int64_t row = tile_base + r;
        for (int c = 0; c < p.n_probe_cols; c++) {
          bool nn = false;
          if (active) {
            p.out_probe[c].data[q] = p.probe[c].data[row];
            nn = tqd::bm_not_null(p.probe[c].bm, row);
          }
          if (p.out_probe[c].bm) {
            const unsigned word = __ballot_sync(0xffffffffu, nn);
            if (lane == 0 && word) {
              if (full_word) p.out_probe[c].bm[q0 >> 5] = word;
              else atomicOr(&p.out_probe[c].bm[q0 >> 5], word);
            }
          }
        }
        for (int c = 0; c < p.n_build_cols; c++) {
          bool nn = false;
          if (active) {
            uint64_t v = 0;
            if (off != OFF_MISS) {
              const uint64_t bpos = (uint64_t)off + j;
              v = p.build[c].data[bpos];
              nn = tqd::bm_not_null(p.build[c].bm, (int64_t)bpos);
            }
            p.out_build[c].data[q] = v;  // defaultInner: NULL (builder.go:463-465)
          }
          if (p.out_build[c].bm) {
            const unsigned word = __ballot_sync(0xffffffffu, nn);
            if (lane == 0 && word) {
              if (full_word) p.out_build[c].bm[q0 >> 5] = word;
              else atomicOr(&p.out_build[c].bm[q0 >> 5], word);
            }
          }
        }
      }
    }
    __syncthreads();  // smem is reused by the next tile
  }
}

// ------------------------------------------------------------------ host side
// Pinned accumulation of ≤1024-row host chunks into one column.
struct HostAccum {
  PinBuf data, bm;
  int64_t n = 0, cap = 0;
  bool has_bm = false;
  int32_t ensure(int64_t rows) {
    if (rows <= cap) return TQ_OK;
    int64_t ncap = cap ? cap : 4096;
    while (ncap < rows) ncap *= 2;
    PinBuf nd, nb;
    TQ_TRY(nd.reserve((size_t)ncap * 8));
    TQ_TRY(nb.reserve(bitmap_alloc_bytes(ncap)));
    if (n) {
      memcpy(nd.p, data.p, (size_t)n * 8);
      memcpy(nb.p, bm.p, bitmap_bytes(n));
    }
    std::swap(data.p, nd.p); std::swap(data.cap, nd.cap);
    std::swap(bm.p, nb.p); std::swap(bm.cap, nb.cap);
    cap = ncap;
    return TQ_OK;
  }
  int32_t append(const tq_column &c, int64_t rows) {
    TQ_TRY(ensure(n + rows));
    memcpy(data.as<uint8_t>() + n * 8, c.data, (size_t)rows * 8);
    host_bitmap_append(bm.as<uint8_t>(), n, c.null_bitmap, rows);
    if (c.null_bitmap) has_bm = true;
    n += rows;
    return TQ_OK;
  }
  void reset() { n = 0; has_bm = false; }
};

struct DevColBuf {
  DevBuf data, bm;
};

struct ResultBatch {
  std::vector<DevColBuf> cols;
  int64_t n = 0;
  uint64_t capacity = 0;
  // host copies (host-consumer path)
  std::vector<PinBuf> h_data, h_bm;
  bool on_host = false;
  cudaEvent_t ev_ready = nullptr;
  ~ResultBatch() { if (ev_ready) cudaEventDestroy(ev_ready); }
};

struct ProbeInputSet {  // device copy of one host probe batch
  std::vector<DevColBuf> cols;
  DevBuf selected;
  cudaEvent_t ev_h2d = nullptr;
  ~ProbeInputSet() { if (ev_h2d) cudaEventDestroy(ev_h2d); }
};

struct PendingBatch {  // a launched probe batch whose row count has not been read back yet
  bool active = false;
  std::unique_ptr<ResultBatch> rb;
  // inputs, kept for a possible re-run on output overflow
  std::vector<DCol> probe;
  const uint8_t *d_selected = nullptr;
  int64_t n = 0;
  bool want_host = false;
  int cursor_slot = 0;
  cudaEvent_t ev_k = nullptr;
};

}  // namespace tq

using namespace tq;

struct tq_join {
  // descriptor
  int join_type = 0, outer_is_right = 0;
  int n_build_cols = 0, n_probe_cols = 0, n_keys = 0;
  int build_types[MAXC], probe_types[MAXC];
  int build_key = 0, probe_key = 0;
  int key_mode = KEYMODE_RAW;
  int64_t batch_rows = 1 << 22;

  enum State { BUILDING, PROBING, CLOSED } state = BUILDING;

  // ---- build side
  std::vector<HostAccum> b_host;          // host chunks accumulate here
  std::vector<std::vector<tq_column>> b_dev_chunks;  // borrowed device chunks
  int build_mem = -1;
  int64_t n_build = 0;
  std::vector<DevColBuf> b_cols;          // materialised inner side (owned) ...
  std::vector<DCol> b_view;               // ... or borrowed view
  std::vector<DevColBuf> csr_cols;        // build columns in CSR order
  DevBuf slots, row_slot, row_ids, counters, worklist, scan_scratch;
  JoinTable table{};
  uint64_t n_slots = 0;
  int64_t n_valid = 0, n_distinct = 0;
  bool build_unique = true;
  int64_t build_ns = 0;

  // ---- probe side
  std::vector<HostAccum> p_host;
  std::vector<uint8_t> p_sel_host;        // selected bytes of the staged chunks (lazily all-ones)
  bool p_sel_any = false;
  ProbeInputSet in_set[2];
  int in_flip = 0;
  PinBuf p_sel_pin[2];
  DevBuf cursors;                         // 2 x {rows, matched} device counters
  PinBuf cursors_host;
  PendingBatch pending;
  std::deque<std::unique_ptr<ResultBatch>> results;
  std::vector<std::unique_ptr<ResultBatch>> free_list;
  std::unique_ptr<ResultBatch> lent;      // batch handed out by next_device
  std::unique_ptr<ResultBatch> host_cur;  // batch being sliced by next()
  int64_t host_cur_pos = 0;
  bool probe_eof = false;
  int64_t probe_rows_total = 0, joined_rows_total = 0, probe_launches = 0, last_probe_ns = 0;
  cudaEvent_t ev_a[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr};  // per cursor slot

  ~tq_join() {
    if (pending.ev_k) cudaEventDestroy(pending.ev_k);
    for (int i = 0; i < 2; i++) {
      if (ev_a[i]) cudaEventDestroy(ev_a[i]);
      if (ev_b[i]) cudaEventDestroy(ev_b[i]);
    }
  }
};

namespace tq {

static int probe_grid(int64_t n) {
  const int64_t tiles = (n + PROBE_TILE - 1) / PROBE_TILE;
  const int64_t cap = (int64_t)rt().sm_count * 6;
  return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}
static int stream_grid(int64_t n) {
  const int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)rt().sm_count * 8;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

static bool type_ok(int t) { return t == TQ_TYPE_INT64 || t == TQ_TYPE_UINT64 || t == TQ_TYPE_FLOAT64; }

static int32_t upload_col(const HostAccum &h, DevColBuf &d, cudaStream_t s) {
  TQ_TRY(d.data.reserve((size_t)(h.n ? h.n : 1) * 8));
  TQ_TRY(d.bm.reserve(bitmap_alloc_bytes(h.n)));
  if (h.n) {
    TQ_CUDA(cudaMemcpyAsync(d.data.p, h.data.p, (size_t)h.n * 8, cudaMemcpyHostToDevice, s));
    TQ_CUDA(cudaMemcpyAsync(d.bm.p, h.bm.p, bitmap_bytes(h.n), cudaMemcpyHostToDevice, s));
  }
  return TQ_OK;
}

static int32_t join_build(tq_join *j) {
  Runtime &r = rt();
  cudaStream_t s = r.compute;
  const int64_t n = j->n_build;
  TQ_CUDA(cudaEventRecord(j->ev_a[0], s));
  // table of distinct keys at load factor <= 0.5
  uint64_t n_slots = 64;
  while (n_slots < (uint64_t)n * 2) n_slots <<= 1;
  if (n_slots > 0xFFFFFFF0ull) { set_error("build side too large: %lld rows", (long long)n); return TQ_ERR_INVALID_ARG; }
  j->n_slots = n_slots;
  TQ_TRY(j->slots.reserve(n_slots * sizeof(Slot)));
  TQ_TRY(j->row_slot.reserve((size_t)(n ? n : 1) * 4));
  TQ_TRY(j->row_ids.reserve((size_t)(n ? n : 1) * 4));
  TQ_TRY(j->counters.reserve(64));
  const uint32_t worklist_cap = (uint32_t)((n / 33) + 2);
  TQ_TRY(j->worklist.reserve((size_t)worklist_cap * 8));
  Slot *slots = j->slots.as<Slot>();
  uint32_t *counters = j->counters.as<uint32_t>();
  TQ_CUDA(cudaMemsetAsync(counters, 0, 64, s));
  k_init_slots<<<stream_grid((int64_t)n_slots), 256, 0, s>>>(slots, n_slots);
  count_launch();
  const DCol key = j->b_view[j->build_key];
  if (n > 0) {
    k_build_insert<<<stream_grid(n), 256, 0, s>>>(key.data, key.bm, n, j->key_mode, slots, n_slots - 1, j->row_slot.as<uint32_t>(), counters);
    count_launch();
  }
  TQ_TRY(check_launch("k_build_insert"));
  // CSR offsets: exclusive scan of slot counts, written into slot.off (AoS stride: 4 words)
  uint64_t *d_total = reinterpret_cast<uint64_t *>(counters + 8);
  TQ_TRY(exclusive_scan_u32(&slots[0].cnt, 4, &slots[0].off, 4, (int64_t)n_slots, d_total, j->scan_scratch, s));
  uint32_t h_counters[16];
  TQ_CUDA(cudaMemcpyAsync(h_counters, counters, 64, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaStreamSynchronize(s));
  const uint64_t total_regular = *reinterpret_cast<uint64_t *>(h_counters + 8);
  const uint32_t sent_cnt = h_counters[0];
  j->n_valid = (int64_t)(total_regular + sent_cnt);
  j->table.slots = slots;
  j->table.mask = n_slots - 1;
  j->table.sent_off = (uint32_t)total_regular;
  j->table.sent_cnt = sent_cnt;
  if (n > 0) {
    k_build_fill<<<stream_grid(n), 256, 0, s>>>(j->row_slot.as<uint32_t>(), n, slots, j->table.sent_off, counters, j->row_ids.as<uint32_t>());
    count_launch();
  }
  k_build_fixsort<<<stream_grid((int64_t)n_slots), 256, 0, s>>>(slots, n_slots, j->row_ids.as<uint32_t>(), counters, j->worklist.as<uint2>(), worklist_cap);
  count_launch();
  TQ_TRY(check_launch("k_build_fixsort"));
  TQ_CUDA(cudaMemcpyAsync(h_counters, counters, 64, cudaMemcpyDeviceToHost, s));
  TQ_CUDA(cudaStreamSynchronize(s));
  uint32_t n_large = h_counters[3];
  j->n_distinct = (int64_t)h_counters[2] + (sent_cnt ? 1 : 0);
  j->build_unique = (j->n_distinct == j->n_valid);
  if (sent_cnt > 1) {  // the sentinel-key segment is sorted like any other large segment
    uint2 w = make_uint2(j->table.sent_off, sent_cnt);
    TQ_CUDA(cudaMemcpyAsync(j->worklist.as<uint2>() + n_large, &w, sizeof(w), cudaMemcpyHostToDevice, s));
    n_large++;
  }
  if (n_large) {
    k_sort_large<<<n_large, 256, 0, s>>>(j->worklist.as<uint2>(), j->row_ids.as<uint32_t>());
    count_launch();
    TQ_TRY(check_launch("k_sort_large"));
  }
  // build columns into CSR order
  j->csr_cols.resize(j->n_build_cols);
  const int64_t nv = j->n_valid;
  for (int c = 0; c < j->n_build_cols; c++) {
    TQ_TRY(j->csr_cols[c].data.reserve((size_t)(nv ? nv : 1) * 8));
    TQ_TRY(j->csr_cols[c].bm.reserve(bitmap_alloc_bytes(nv)));
    if (nv) {
      k_gather_col<<<stream_grid(nv), 256, 0, s>>>(j->b_view[c].data, j->b_view[c].bm, j->row_ids.as<uint32_t>(), nv,
                                                    j->csr_cols[c].data.as<uint64_t>(), j->csr_cols[c].bm.as<uint32_t>());
      count_launch();
    }
  }
  TQ_TRY(check_launch("k_gather_col"));
  TQ_CUDA(cudaEventRecord(j->ev_b[0], s));
  TQ_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  TQ_CUDA(cudaEventElapsedTime(&ms, j->ev_a[0], j->ev_b[0]));
  j->build_ns = (int64_t)(ms * 1e6);
  // the row-order copies are no longer needed
  j->row_slot.release();
  j->row_ids.release();
  j->b_cols.clear();
  j->b_view.clear();
  return TQ_OK;
}

static std::unique_ptr<ResultBatch> get_result_batch(tq_join *j) {
  std::unique_ptr<ResultBatch> rb;
  if (!j->free_list.empty()) { rb = std::move(j->free_list.back()); j->free_list.pop_back(); }
  else rb.reset(new ResultBatch());
  rb->n = 0;
  rb->on_host = false;
  return rb;
}

// Enqueue one probe launch for `n` rows of device columns `probe` into rb (capacity rows).
static int32_t launch_probe(tq_join *j, const std::vector<DCol> &probe, const uint8_t *d_selected, int64_t n, ResultBatch *rb, uint64_t capacity,
                            int cursor_slot) {
  Runtime &r = rt();
  cudaStream_t s = r.compute;
  const int ncols = j->n_build_cols + j->n_probe_cols;
  rb->cols.resize(ncols);
  rb->capacity = capacity;
  const int build_base = j->outer_is_right ? 0 : j->n_probe_cols;   // joiner.go:145-150: lhs ++ rhs
  const int probe_base = j->outer_is_right ? j->n_build_cols : 0;
  ProbeParams p{};
  p.n_probe_cols = j->n_probe_cols;
  p.n_build_cols = j->n_build_cols;
  p.selected = d_selected;
  p.key_col = j->probe_key;
  p.key_mode = j->key_mode;
  p.is_outer = (j->join_type != TQ_JOIN_INNER);
  p.n = n;
  p.capacity = capacity;
  unsigned long long *cur = j->cursors.as<unsigned long long>() + 2 * cursor_slot;
  p.cursor = cur;
  TQ_CUDA(cudaMemsetAsync(cur, 0, 16, s));
  for (int c = 0; c < ncols; c++) {
    TQ_TRY(rb->cols[c].data.reserve((size_t)(capacity ? capacity : 1) * 8));
    TQ_TRY(rb->cols[c].bm.reserve(bitmap_alloc_bytes((int64_t)capacity)));
  }
  for (int c = 0; c < j->n_probe_cols; c++) {
    p.probe[c] = probe[c];
    DevColBuf &o = rb->cols[probe_base + c];
    p.out_probe[c].data = o.data.as<uint64_t>();
    const bool may_null = probe[c].bm != nullptr;
    TQ_CUDA(cudaMemsetAsync(o.bm.p, may_null ? 0x00 : 0xFF, bitmap_alloc_bytes((int64_t)capacity), s));
    p.out_probe[c].bm = may_null ? o.bm.as<uint32_t>() : nullptr;
  }
  for (int c = 0; c < j->n_build_cols; c++) {
    p.build[c].data = j->csr_cols[c].data.as<uint64_t>();
    p.build[c].bm = j->csr_cols[c].bm.as<uint32_t>();
    DevColBuf &o = rb->cols[build_base + c];
    p.out_build[c].data = o.data.as<uint64_t>();
    TQ_CUDA(cudaMemsetAsync(o.bm.p, 0x00, bitmap_alloc_bytes((int64_t)capacity), s));
    p.out_build[c].bm = o.bm.as<uint32_t>();
  }
  TQ_CUDA(cudaEventRecord(j->ev_a[cursor_slot], s));
  k_probe<<<probe_grid(n), PROBE_THREADS, 0, s>>>(p, j->table);
  count_launch();
  j->probe_launches++;
  TQ_TRY(check_launch("k_probe"));
  TQ_CUDA(cudaEventRecord(j->ev_b[cursor_slot], s));
  TQ_CUDA(cudaMemcpyAsync(j->cursors_host.as<unsigned long long>() + 2 * cursor_slot, cur, 16, cudaMemcpyDeviceToHost, s));
  return TQ_OK;
}

static int32_t enqueue_d2h(tq_join *j, ResultBatch *rb) {
  Runtime &r = rt();
  const int ncols = (int)rb->cols.size();
  rb->h_data.resize(ncols);
  rb->h_bm.resize(ncols);
  if (!rb->ev_ready) TQ_CUDA(cudaEventCreateWithFlags(&rb->ev_ready, cudaEventDisableTiming));
  for (int c = 0; c < ncols; c++) {
    TQ_TRY(rb->h_data[c].reserve((size_t)(rb->n ? rb->n : 1) * 8));
    TQ_TRY(rb->h_bm[c].reserve(bitmap_alloc_bytes(rb->n)));
    if (rb->n) {
      TQ_CUDA(cudaMemcpyAsync(rb->h_data[c].p, rb->cols[c].data.p, (size_t)rb->n * 8, cudaMemcpyDeviceToHost, r.d2h));
      TQ_CUDA(cudaMemcpyAsync(rb->h_bm[c].p, rb->cols[c].bm.p, bitmap_bytes(rb->n), cudaMemcpyDeviceToHost, r.d2h));
    }
  }
  TQ_CUDA(cudaEventRecord(rb->ev_ready, r.d2h));
  rb->on_host = true;
  return TQ_OK;
}

// Wait for the pending batch, re-run it if the output did not fit, queue its result.
static int32_t finalize_pending(tq_join *j) {
  PendingBatch &pb = j->pending;
  if (!pb.active) return TQ_OK;
  Runtime &r = rt();
  TQ_CUDA(cudaEventSynchronize(pb.ev_k));
  unsigned long long *hc = j->cursors_host.as<unsigned long long>() + 2 * pb.cursor_slot;
  uint64_t produced = hc[0];
  float ms = 0;
  if (cudaEventElapsedTime(&ms, j->ev_a[pb.cursor_slot], j->ev_b[pb.cursor_slot]) == cudaSuccess) j->last_probe_ns = (int64_t)(ms * 1e6);
  else cudaGetLastError();
  if (produced > pb.rb->capacity) {
    // duplicate build keys: the first launch served as the count pass; run again with the exact size
    TQ_TRY(launch_probe(j, pb.probe, pb.d_selected, pb.n, pb.rb.get(), produced, pb.cursor_slot));
    TQ_CUDA(cudaStreamSynchronize(r.compute));
    produced = hc[0];
    if (produced > pb.rb->capacity) { set_error("join output size changed between passes"); return TQ_ERR_CUDA; }
    if (cudaEventElapsedTime(&ms, j->ev_a[pb.cursor_slot], j->ev_b[pb.cursor_slot]) == cudaSuccess) j->last_probe_ns = (int64_t)(ms * 1e6);
  }
  pb.rb->n = (int64_t)produced;
  j->joined_rows_total += (int64_t)produced;
  if (pb.want_host) {
    TQ_CUDA(cudaStreamWaitEvent(r.d2h, pb.ev_k, 0));
    TQ_TRY(enqueue_d2h(j, pb.rb.get()));
  }
  j->results.push_back(std::move(pb.rb));
  pb.active = false;
  return TQ_OK;
}

// Start a probe batch on device columns in cursor/input slot `slot`; the PREVIOUS batch (other slot) is
// finalised after this one is enqueued, so the host-side wait overlaps GPU work.
static int32_t start_batch(tq_join *j, const std::vector<DCol> &probe, const uint8_t *d_selected, int64_t n, bool want_host, int slot) {
  if (n == 0) return TQ_OK;
  Runtime &r = rt();
  if (j->pending.active && j->pending.cursor_slot == slot) TQ_TRY(finalize_pending(j));
  // a join on unique build keys produces at most one row per probe row; with duplicate keys the
  // first launch doubles as the count pass (finalize_pending re-runs with the exact size)
  const uint64_t capacity = (uint64_t)n;
  std::unique_ptr<ResultBatch> rb = get_result_batch(j);
  TQ_TRY(launch_probe(j, probe, d_selected, n, rb.get(), capacity, slot));
  cudaEvent_t ev = nullptr;
  TQ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  TQ_CUDA(cudaEventRecord(ev, r.compute));
  TQ_TRY(finalize_pending(j));
  PendingBatch &pb = j->pending;
  if (pb.ev_k) cudaEventDestroy(pb.ev_k);
  pb.ev_k = ev;
  pb.active = true;
  pb.rb = std::move(rb);
  pb.probe = probe;
  pb.d_selected = d_selected;
  pb.n = n;
  pb.want_host = want_host;
  pb.cursor_slot = slot;
  j->probe_rows_total += n;
  return TQ_OK;
}

// Ship one piece of host rows to the device (double-buffered input sets) and start its probe.
static int32_t process_host_piece(tq_join *j, const tq_column *cols, int64_t row0, int64_t rows, const uint8_t *selected) {
  Runtime &r = rt();
  const int slot = j->in_flip;
  j->in_flip ^= 1;
  // the batch that last read this input set must be complete before the set is overwritten
  if (j->pending.active && j->pending.cursor_slot == slot) TQ_TRY(finalize_pending(j));
  ProbeInputSet &in = j->in_set[slot];
  if (!in.ev_h2d) TQ_CUDA(cudaEventCreateWithFlags(&in.ev_h2d, cudaEventDisableTiming));
  in.cols.resize(j->n_probe_cols);
  std::vector<DCol> view(j->n_probe_cols);
  if ((row0 & 7) != 0) { set_error("internal: unaligned host piece"); return TQ_ERR_INVALID_ARG; }
  for (int c = 0; c < j->n_probe_cols; c++) {
    TQ_TRY(in.cols[c].data.reserve((size_t)rows * 8));
    TQ_CUDA(cudaMemcpyAsync(in.cols[c].data.p, cols[c].data + row0 * 8, (size_t)rows * 8, cudaMemcpyHostToDevice, r.h2d));
    view[c].data = in.cols[c].data.as<uint64_t>();
    view[c].bm = nullptr;
    if (cols[c].null_bitmap) {
      TQ_TRY(in.cols[c].bm.reserve(bitmap_alloc_bytes(rows)));
      TQ_CUDA(cudaMemcpyAsync(in.cols[c].bm.p, cols[c].null_bitmap + (row0 >> 3), bitmap_bytes(rows), cudaMemcpyHostToDevice, r.h2d));
      view[c].bm = in.cols[c].bm.as<uint32_t>();
    }
  }
  const uint8_t *d_sel = nullptr;
  if (selected) {
    TQ_TRY(in.selected.reserve((size_t)rows));
    TQ_TRY(j->p_sel_pin[slot].reserve((size_t)rows));
    memcpy(j->p_sel_pin[slot].p, selected + row0, (size_t)rows);
    TQ_CUDA(cudaMemcpyAsync(in.selected.p, j->p_sel_pin[slot].p, (size_t)rows, cudaMemcpyHostToDevice, r.h2d));
    d_sel = in.selected.as<uint8_t>();
  }
  TQ_CUDA(cudaEventRecord(in.ev_h2d, r.h2d));
  TQ_CUDA(cudaStreamWaitEvent(r.compute, in.ev_h2d, 0));
  return start_batch(j, view, d_sel, rows, /*want_host=*/true, slot);
}

static int32_t flush_probe_staging(tq_join *j) {
  if (j->p_host.empty() || j->p_host[0].n == 0) return TQ_OK;
  const int64_t rows = j->p_host[0].n;
  std::vector<tq_column> cols(j->n_probe_cols);
  for (int c = 0; c < j->n_probe_cols; c++) {
    cols[c].length = rows;
    cols[c].data = j->p_host[c].data.as<uint8_t>();
    cols[c].null_bitmap = j->p_host[c].has_bm ? j->p_host[c].bm.as<uint8_t>() : nullptr;
    cols[c].offsets = nullptr;
  }
  const uint8_t *sel = j->p_sel_any ? j->p_sel_host.data() : nullptr;
  // The staging buffers are reused right after this call, so the H2D copies must have completed.
  int32_t st = process_host_piece(j, cols.data(), 0, rows, sel);
  if (st == TQ_OK) {
    cudaError_t e = cudaStreamSynchronize(rt().h2d);
    if (e != cudaSuccess) st = cuda_fail(e, "sync h2d", __FILE__, __LINE__);
  }
  for (auto &h : j->p_host) h.reset();
  j->p_sel_host.clear();
  j->p_sel_any = false;
  return st;
}

static void recycle(tq_join *j, std::unique_ptr<ResultBatch> rb) {
  if (rb) j->free_list.push_back(std::move(rb));
}

}  // namespace tq

extern "C" {

int32_t tq_join_create(const tq_join_desc *d, tq_join **out) {
  if (!d || !out) return TQ_ERR_INVALID_ARG;
  *out = nullptr;
  TQ_TRY(ensure_init());
  if (d->join_type < TQ_JOIN_INNER || d->join_type > TQ_JOIN_RIGHT_OUTER) { set_error("unsupported join type %d", d->join_type); return TQ_ERR_INVALID_ARG; }
  if (d->n_build_cols < 1 || d->n_build_cols > MAXC || d->n_probe_cols < 1 || d->n_probe_cols > MAXC) {
    set_error("join sides must have 1..%d columns", MAXC);
    return TQ_ERR_INVALID_ARG;
  }
  if (d->n_keys != 1) {
    set_error("hash join on %d key columns: only single-column keys are implemented", d->n_keys);
    return TQ_ERR_UNSUPPORTED_TYPE;
  }
  for (int c = 0; c < d->n_build_cols; c++)
    if (!type_ok(d->build_types[c])) { set_error("unsupport column type for encode %d", d->build_types[c]); return TQ_ERR_UNSUPPORTED_TYPE; }
  for (int c = 0; c < d->n_probe_cols; c++)
    if (!type_ok(d->probe_types[c])) { set_error("unsupport column type for encode %d", d->probe_types[c]); return TQ_ERR_UNSUPPORTED_TYPE; }
  if (d->build_key_idx[0] < 0 || d->build_key_idx[0] >= d->n_build_cols || d->probe_key_idx[0] < 0 || d->probe_key_idx[0] >= d->n_probe_cols)
    return TQ_ERR_INVALID_ARG;
  // LeftOuter keeps the left child as the outer side, RightOuter the right child (builder.go:451-477)
  if (d->join_type == TQ_JOIN_LEFT_OUTER && d->outer_is_right) { set_error("left outer join needs outer_is_right == 0"); return TQ_ERR_INVALID_ARG; }
  if (d->join_type == TQ_JOIN_RIGHT_OUTER && !d->outer_is_right) { set_error("right outer join needs outer_is_right == 1"); return TQ_ERR_INVALID_ARG; }
  tq_join *j = new (std::nothrow) tq_join();
  if (!j) return TQ_ERR_OOM;
  j->join_type = d->join_type;
  j->outer_is_right = d->outer_is_right ? 1 : 0;
  j->n_build_cols = d->n_build_cols;
  j->n_probe_cols = d->n_probe_cols;
  j->n_keys = 1;
  for (int c = 0; c < d->n_build_cols; c++) j->build_types[c] = d->build_types[c];
  for (int c = 0; c < d->n_probe_cols; c++) j->probe_types[c] = d->probe_types[c];
  j->build_key = d->build_key_idx[0];
  j->probe_key = d->probe_key_idx[0];
  const int bt = j->build_types[j->build_key], pt = j->probe_types[j->probe_key];
  const bool bf = bt == TQ_TYPE_FLOAT64, pf = pt == TQ_TYPE_FLOAT64;
  if (bf != pf) j->key_mode = KEYMODE_NEVER;
  else if (!bf && bt != pt) j->key_mode = KEYMODE_NO_SIGNBIT;
  else j->key_mode = KEYMODE_RAW;
  if (d->probe_batch_rows > 0) j->batch_rows = (d->probe_batch_rows + 63) & ~63ll;
  j->b_host.resize(j->n_build_cols);
  j->p_host.resize(j->n_probe_cols);
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; i++) {
    e = cudaEventCreate(&j->ev_a[i]);
    if (e == cudaSuccess) e = cudaEventCreate(&j->ev_b[i]);
  }
  if (e != cudaSuccess) { delete j; return cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); }
  int32_t st = j->cursors.reserve(64);
  if (st == TQ_OK) st = j->cursors_host.reserve(64);
  if (st != TQ_OK) { delete j; return st; }
  *out = j;
  return TQ_OK;
}

int32_t tq_join_put_build(tq_join *j, const tq_column *cols, int32_t mem) {
  if (!j || !cols) return TQ_ERR_INVALID_ARG;
  TQ_TRY(ensure_init());
  if (j->state != tq_join::BUILDING) { set_error("put_build after finalize_build"); return TQ_ERR_STATE; }
  if (j->build_mem >= 0 && j->build_mem != mem) { set_error("build chunks must all be host or all device"); return TQ_ERR_INVALID_ARG; }
  j->build_mem = mem;
  const int64_t rows = cols[0].length;
  if (rows < 0) return TQ_ERR_INVALID_ARG;
  for (int c = 0; c < j->n_build_cols; c++) {
    if (cols[c].length != rows) { set_error("ragged build chunk"); return TQ_ERR_INVALID_ARG; }
    if (cols[c].offsets) { set_error("unsupport column type for encode (var-len column %d)", c); return TQ_ERR_UNSUPPORTED_TYPE; }
    if (rows && !cols[c].data) return TQ_ERR_INVALID_ARG;
  }
  if (rows == 0) return TQ_OK;
  if (mem == TQ_MEM_HOST) {
    for (int c = 0; c < j->n_build_cols; c++) TQ_TRY(j->b_host[c].append(cols[c], rows));
  } else {
    j->b_dev_chunks.emplace_back(cols, cols + j->n_build_cols);
  }
  j->n_build += rows;
  return TQ_OK;
}

int32_t tq_join_finalize_build(tq_join *j) {
  if (!j) return TQ_ERR_INVALID_ARG;
  TQ_TRY(ensure_init());
  if (j->state != tq_join::BUILDING) { set_error("finalize_build called twice"); return TQ_ERR_STATE; }
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  j->b_view.assign(j->n_build_cols, DCol());
  if (j->build_mem == TQ_MEM_DEVICE && j->b_dev_chunks.size() == 1) {
    for (int c = 0; c < j->n_build_cols; c++) {
      j->b_view[c].data = (const uint64_t *)j->b_dev_chunks[0][c].data;
      j->b_view[c].bm = (const uint32_t *)j->b_dev_chunks[0][c].null_bitmap;
    }
  } else if (j->build_mem == TQ_MEM_DEVICE) {
    // several device chunks: concatenate (data D2D; bitmaps need 8-row alignment at chunk boundaries)
    j->b_cols.resize(j->n_build_cols);
    for (int c = 0; c < j->n_build_cols; c++) {
      TQ_TRY(j->b_cols[c].data.reserve((size_t)j->n_build * 8));
      TQ_TRY(j->b_cols[c].bm.reserve(bitmap_alloc_bytes(j->n_build)));
      int64_t off = 0;
      bool any_bm = false;
      for (auto &ch : j->b_dev_chunks) any_bm |= (ch[c].null_bitmap != nullptr);
      if (any_bm) TQ_CUDA(cudaMemsetAsync(j->b_cols[c].bm.p, 0xFF, bitmap_alloc_bytes(j->n_build), r.compute));
      for (auto &ch : j->b_dev_chunks) {
        const int64_t rows = ch[c].length;
        TQ_CUDA(cudaMemcpyAsync(j->b_cols[c].data.as<uint8_t>() + off * 8, ch[c].data, (size_t)rows * 8, cudaMemcpyDeviceToDevice, r.compute));
        if (ch[c].null_bitmap) {
          if (off & 7) { set_error("device build chunks with NULL bitmaps must have row counts that are multiples of 8"); return TQ_ERR_INVALID_ARG; }
          TQ_CUDA(cudaMemcpyAsync(j->b_cols[c].bm.as<uint8_t>() + (off >> 3), ch[c].null_bitmap, bitmap_bytes(rows), cudaMemcpyDeviceToDevice, r.compute));
        }
        off += rows;
      }
      j->b_view[c].data = j->b_cols[c].data.as<uint64_t>();
      j->b_view[c].bm = any_bm ? j->b_cols[c].bm.as<uint32_t>() : nullptr;
    }
  } else {
    j->b_cols.resize(j->n_build_cols);
    for (int c = 0; c < j->n_build_cols; c++) {
      TQ_TRY(upload_col(j->b_host[c], j->b_cols[c], r.compute));
      j->b_view[c].data = j->b_cols[c].data.as<uint64_t>();
      j->b_view[c].bm = j->b_host[c].has_bm ? j->b_cols[c].bm.as<uint32_t>() : nullptr;
    }
  }
  TQ_TRY(join_build(j));
  for (auto &h : j->b_host) { h.data.release(); h.bm.release(); }
  j->b_dev_chunks.clear();
  j->state = tq_join::PROBING;
  return TQ_OK;
}

int32_t tq_join_put_probe(tq_join *j, const tq_column *cols, const uint8_t *selected, int32_t mem) {
  if (!j || !cols) return TQ_ERR_INVALID_ARG;
  TQ_TRY(ensure_init());
  if (j->state != tq_join::PROBING) { set_error("put_probe before finalize_build"); return TQ_ERR_STATE; }
  if (j->probe_eof) { set_error("put_probe after probe_eof"); return TQ_ERR_STATE; }
  const int64_t rows = cols[0].length;
  if (rows < 0) return TQ_ERR_INVALID_ARG;
  for (int c = 0; c < j->n_probe_cols; c++) {
    if (cols[c].length != rows) { set_error("ragged probe chunk"); return TQ_ERR_INVALID_ARG; }
    if (cols[c].offsets) { set_error("unsupport column type for encode (var-len column %d)", c); return TQ_ERR_UNSUPPORTED_TYPE; }
    if (rows && !cols[c].data) return TQ_ERR_INVALID_ARG;
  }
  if (rows == 0) return TQ_OK;
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  if (mem == TQ_MEM_DEVICE) {
    TQ_TRY(flush_probe_staging(j));
    std::vector<DCol> view(j->n_probe_cols);
    for (int c = 0; c < j->n_probe_cols; c++) {
      view[c].data = (const uint64_t *)cols[c].data;
      view[c].bm = (const uint32_t *)cols[c].null_bitmap;
    }
    const int slot = j->pending.active ? 1 - j->pending.cursor_slot : 0;
    const uint8_t *d_sel = nullptr;
    if (selected) {  // `selected` is host memory by contract
      ProbeInputSet &in = j->in_set[slot];
      TQ_TRY(in.selected.reserve((size_t)rows));
      TQ_CUDA(cudaMemcpyAsync(in.selected.p, selected, (size_t)rows, cudaMemcpyHostToDevice, r.compute));
      TQ_CUDA(cudaStreamSynchronize(r.compute));
      d_sel = in.selected.as<uint8_t>();
    }
    return start_batch(j, view, d_sel, rows, /*want_host=*/false, slot);
  }
  if (rows >= j->batch_rows) {
    // a large host column (not the ≤1024-row chunk protocol): stream it in batch-sized pieces straight from the caller's buffer
    TQ_TRY(flush_probe_staging(j));
    for (int64_t row0 = 0; row0 < rows; row0 += j->batch_rows) {
      const int64_t piece = rows - row0 < j->batch_rows ? rows - row0 : j->batch_rows;
      TQ_TRY(process_host_piece(j, cols, row0, piece, selected));
    }
    TQ_CUDA(cudaStreamSynchronize(r.h2d));  // caller may reuse its buffers on return
    return TQ_OK;
  }
  if (j->p_host[0].n + rows > j->batch_rows) TQ_TRY(flush_probe_staging(j));
  const int64_t before = j->p_host[0].n;
  for (int c = 0; c < j->n_probe_cols; c++) TQ_TRY(j->p_host[c].append(cols[c], rows));
  if (selected && !j->p_sel_any) { j->p_sel_host.assign((size_t)before, 1); j->p_sel_any = true; }
  if (j->p_sel_any) {
    if (selected) j->p_sel_host.insert(j->p_sel_host.end(), selected, selected + rows);
    else j->p_sel_host.insert(j->p_sel_host.end(), (size_t)rows, 1);
  }
  return TQ_OK;
}

int32_t tq_join_probe_eof(tq_join *j) {
  if (!j) return TQ_ERR_INVALID_ARG;
  TQ_TRY(ensure_init());
  if (j->state != tq_join::PROBING) { set_error("probe_eof before finalize_build"); return TQ_ERR_STATE; }
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  TQ_TRY(flush_probe_staging(j));
  TQ_TRY(finalize_pending(j));
  j->probe_eof = true;
  return TQ_OK;
}

int32_t tq_join_next(tq_join *j, int64_t max_rows, tq_column *out_cols, int64_t *n_rows, int32_t *eof) {
  if (!j || !out_cols || !n_rows || !eof || max_rows <= 0) return TQ_ERR_INVALID_ARG;
  TQ_TRY(ensure_init());
  *n_rows = 0;
  *eof = 0;
  if (j->state == tq_join::BUILDING) { set_error("next before finalize_build"); return TQ_ERR_STATE; }
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  const int ncols = j->n_build_cols + j->n_probe_cols;
  for (;;) {
    if (j->host_cur && j->host_cur_pos < j->host_cur->n) break;
    if (j->host_cur) { recycle(j, std::move(j->host_cur)); j->host_cur_pos = 0; }
    if (j->results.empty()) TQ_TRY(finalize_pending(j));
    if (j->results.empty()) {
      *eof = j->probe_eof ? 1 : 0;
      for (int c = 0; c < ncols; c++) out_cols[c].length = 0;
      return TQ_OK;
    }
    j->host_cur = std::move(j->results.front());
    j->results.pop_front();
    j->host_cur_pos = 0;
    if (!j->host_cur->on_host) {
      TQ_CUDA(cudaStreamSynchronize(r.compute));
      TQ_TRY(enqueue_d2h(j, j->host_cur.get()));
    }
    TQ_CUDA(cudaEventSynchronize(j->host_cur->ev_ready));
  }
  ResultBatch *rb = j->host_cur.get();
  const int64_t take = (rb->n - j->host_cur_pos) < max_rows ? (rb->n - j->host_cur_pos) : max_rows;
  for (int c = 0; c < ncols; c++) {
    if (!out_cols[c].data || !out_cols[c].null_bitmap) { set_error("output column %d needs data and null_bitmap buffers", c); return TQ_ERR_INVALID_ARG; }
    memcpy(out_cols[c].data, rb->h_data[c].as<uint8_t>() + j->host_cur_pos * 8, (size_t)take * 8);
    host_bitmap_extract(out_cols[c].null_bitmap, rb->h_bm[c].as<uint8_t>(), j->host_cur_pos, take);
    out_cols[c].length = take;
  }
  j->host_cur_pos += take;
  *n_rows = take;
  return TQ_OK;
}

int32_t tq_join_next_device(tq_join *j, tq_column *out_cols, int64_t *n_rows, int32_t *eof) {
  if (!j || !out_cols || !n_rows || !eof) return TQ_ERR_INVALID_ARG;
  TQ_TRY(ensure_init());
  *n_rows = 0;
  *eof = 0;
  if (j->state == tq_join::BUILDING) { set_error("next before finalize_build"); return TQ_ERR_STATE; }
  Runtime &r = rt();
  std::lock_guard<std::recursive_mutex> lk(r.mu);
  recycle(j, std::move(j->lent));
  if (j->results.empty()) TQ_TRY(finalize_pending(j));
  if (j->results.empty()) {
    *eof = j->probe_eof ? 1 : 0;
    return TQ_OK;
  }
  j->lent = std::move(j->results.front());
  j->results.pop_front();
  const int ncols = j->n_build_cols + j->n_probe_cols;
  for (int c = 0; c < ncols; c++) {
    out_cols[c].length = j->lent->n;
    out_cols[c].data = j->lent->cols[c].data.as<uint8_t>();
    out_cols[c].null_bitmap = j->lent->cols[c].bm.as<uint8_t>();
    out_cols[c].offsets = nullptr;
  }
  *n_rows = j->lent->n;
  return TQ_OK;
}

int32_t tq_join_stats(tq_join *j, int64_t *s) {
  if (!j || !s) return TQ_ERR_INVALID_ARG;
  s[0] = j->n_valid;
  s[1] = j->n_distinct;
  s[2] = 1;
  s[3] = j->probe_rows_total;
  s[4] = j->joined_rows_total;
  s[5] = j->last_probe_ns;
  s[6] = j->build_ns;
  s[7] = j->probe_launches;
  return TQ_OK;
}

int32_t tq_join_destroy(tq_join *j) {
  if (!j) return TQ_OK;
  if (rt().inited) {
    cudaSetDevice(rt().device);
    cudaDeviceSynchronize();  // Close may arrive with work in flight (join_test.go:175-182 early Close)
  }
  delete j;
  return TQ_OK;
}

}  // extern "C"
