// join_stream.cuh — the streaming PK-FK pipeline of HashJoinExec (included by join.cu, which defines ScatterParams,
// ProbeParams, JoinTable, ld_pair and the TMA / mbarrier helpers it uses).
//
// Three kernels, all HBM/L2-bound integer work (no tensor cores: nothing here is a contraction):
//
//   k_scatter_aos<NC>   radix scatter of NC 8-byte columns into array-of-structs partition slabs.  Input tiles arrive in
//                       shared memory by TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) one tile ahead; the rank of
//                       a row inside its partition comes from warp ballots over the partition-id bits plus a per-(warp, bin)
//                       counter that only the bin's leader lane touches — no shared-memory atomic per row; rows leave through
//                       an AoS staging area as full-sector 16-byte stores.
//   k_probe_pos<NP,NB>  probe of one partition against its L2-resident table.  Slab tiles are TMA-streamed through a ring of
//                       shared-memory stages by a producer warp; every consumer warp is independent (no CTA barrier, no
//                       output cursor): the output position of a probe row is its position in the partition order, so a
//                       warp's 32 rows go to 32 consecutive slots of every output column, and a ballot word records which
//                       slots are real.  The holes (misses, and the padding of each partition to a multiple of 32) are
//                       filled afterwards from the tail of the result (k_hole_*): for a foreign-key join that is a few
//                       thousand rows.  The contract is the result MULTISET; order is not.
//   k_build_part<NB>    build of the partition tables from build-side AoS slabs: one CTA initialises a partition's table
//                       and inserts its rows while the table is L2-resident (the global insert touched a random DRAM line
//                       per row).
#pragma once

namespace tq {

__device__ __forceinline__ bool mbar_try_wait(uint64_t *mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(mbar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_spin(uint64_t *mbar, uint32_t parity) {
  while (!mbar_try_wait(mbar, parity)) {}
}
__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ AoS scatter
static constexpr int SA_WARPS = 16;
static constexpr int SA_THREADS = SA_WARPS * 32;
static constexpr int SA_MAX_PBITS = 9;                       // 512 partitions (+ outer bin + trash bin) keep the counter matrix at 16 KB
static constexpr int SA_MAX_BINS = (1 << SA_MAX_PBITS) + 2;

struct ScatterAosParams {
  ScatterParams sp;        // columns, key, selected, pbits, cursors, slab limits (sp.out is unused)
  uint64_t *out;           // AoS slabs: row r = out[r * NC .. r * NC + NC)
  int use_tma;             // every input column is 16-byte aligned: whole tiles are fetched by TMA bulk copies
};

template <int T>
__host__ __device__ constexpr int sa_smem_bytes(int nc, int bins) {
  // 2 input stages + AoS staging + counters + bin tables + sorted position -> bin + 2 mbarriers
  return 2 * nc * T * 8 + nc * T * 8 + ((SA_WARPS * bins * 2 + 15) & ~15) + 3 * ((bins * 4 + 15) & ~15) + T * 2 + 64;
}

template <int NC, int T>
__global__ void __launch_bounds__(SA_THREADS, (sa_smem_bytes<T>(NC, 130) <= 110 * 1024 ? 2 : 1)) k_scatter_aos(const ScatterAosParams q) {
  constexpr int R = T / SA_THREADS;  // rows per thread per tile
  extern __shared__ __align__(128) unsigned char s_raw[];
  const ScatterParams &p = q.sp;
  const int n_part_bins = scatter_bins(p);      // partitions + the outer-join bin
  const int n_bins = n_part_bins + 1;           // + trash (rows that produce nothing)
  const int trash = n_bins - 1;
  int nbits = 1;
  while ((1 << nbits) < n_bins) nbits++;
  uint64_t *s_in = reinterpret_cast<uint64_t *>(s_raw);                       // [2][NC][T]
  uint64_t *s_sorted = s_in + 2 * NC * T;                                      // [T][NC]
  uint16_t *s_cnt = reinterpret_cast<uint16_t *>(s_sorted + NC * T);           // [SA_WARPS][n_bins]
  uint32_t *s_start = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(s_cnt) + ((SA_WARPS * n_bins * 2 + 15) & ~15));
  uint32_t *s_gdelta = s_start + ((n_bins + 3) & ~3);
  uint32_t *s_imax = s_gdelta + ((n_bins + 3) & ~3);
  uint16_t *s_spid = reinterpret_cast<uint16_t *>(s_imax + ((n_bins + 3) & ~3));  // [T]
  uint64_t *s_mbar = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(s_spid + T) + 15) & ~(uintptr_t)15);  // [2]
  __shared__ uint32_t s_warp[SA_WARPS + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1;
  const int kc = p.key_col;
  const int64_t n_tiles = (p.n + T - 1) / T;
  const int bpt = (n_bins + SA_THREADS - 1) / SA_THREADS;

  if (tid == 0) {
    mbar_init(&s_mbar[0], 1);
    mbar_init(&s_mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto tile_is_tma = [&](int64_t tile) { return q.use_tma && (tile + 1) * (int64_t)T <= p.n; };
  auto issue = [&](int64_t tile, int stage) {  // thread 0 only
    mbar_expect_tx(&s_mbar[stage], (uint32_t)(NC * T * 8));
#pragma unroll
    for (int c = 0; c < NC; c++) tma_load_1d(s_in + (stage * NC + c) * T, p.in[c].data + tile * T, (uint32_t)(T * 8), &s_mbar[stage]);
  };
  if (tid == 0) {
    const int64_t t0 = blockIdx.x, t1 = t0 + gridDim.x;
    if (t0 < n_tiles && tile_is_tma(t0)) issue(t0, 0);
    if (t1 < n_tiles && tile_is_tma(t1)) issue(t1, 1);
  }
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
    const int stage = it & 1;
    const int64_t tile_base = tile * T;
    const int rows = (int)((p.n - tile_base) < T ? (p.n - tile_base) : T);
    uint64_t *in_s = s_in + stage * NC * T;
    for (int i = tid; i < SA_WARPS * n_bins; i += SA_THREADS) s_cnt[i] = 0;
    if (tile_is_tma(tile)) {
      mbar_spin(&s_mbar[stage], (uint32_t)((it >> 1) & 1));
    } else {  // unaligned caller buffers or the ragged last tile: plain loads
#pragma unroll
      for (int c = 0; c < NC; c++)
        for (int i = tid; i < rows; i += SA_THREADS) in_s[c * T + i] = tqd::ld_stream_u64(p.in[c].data + tile_base + i);
    }
    __syncthreads();
    // ---- rank every row inside its bin: ballots over the bin-id bits give the lanes of this warp-step that share the bin;
    // the lowest of them bumps the (warp, bin) counter for all of them
    uint16_t pid[R], rank[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      const int row = k * SA_THREADS + tid;
      uint32_t b = (uint32_t)trash;
      if (row < rows) {
        const uint64_t key = in_s[kc * T + row];
        const bool sel = p.selected ? (p.selected[tile_base + row] != 0) : true;
        if (sel && key_valid(key, true, p.key_mode)) b = scatter_pid(p, key);
        else if (p.is_outer) b = (uint32_t)(n_part_bins - 1);
      }
      unsigned peers = 0xffffffffu;
      for (int bit = 0; bit < nbits; bit++) {
        const bool one = (b >> bit) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, one);
        peers &= one ? vote : ~vote;
      }
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = s_cnt[warp * n_bins + b];
        s_cnt[warp * n_bins + b] = (uint16_t)(old + __popc(peers));
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      pid[k] = (uint16_t)b;
      rank[k] = (uint16_t)(old + __popc(peers & lt_mask));
      __syncwarp();
    }
    __syncthreads();
    // ---- per bin: exclusive prefix over the warps (in place), tile total; then an exclusive scan over the bins
    uint32_t tsum = 0;
    for (int j = 0; j < bpt; j++) {
      const int b = tid * bpt + j;
      if (b < n_bins) {
        uint32_t run = 0;
        for (int w = 0; w < SA_WARPS; w++) {
          const uint32_t c = s_cnt[w * n_bins + b];
          s_cnt[w * n_bins + b] = (uint16_t)run;
          run += c;
        }
        s_start[b] = run;  // the bin's tile total for now
        tsum += run;
      }
    }
    uint32_t inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += x; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = (lane < SA_WARPS) ? s_warp[lane] : 0, winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += x; }
      if (lane < SA_WARPS) s_warp[lane] = winc - w;
    }
    __syncthreads();
    uint32_t run = inc - tsum + s_warp[warp];
    for (int j = 0; j < bpt; j++) {
      const int b = tid * bpt + j;
      if (b < n_bins) {
        const uint32_t c = s_start[b];
        s_start[b] = run;
        if (c && b != trash) {
          const uint32_t g = atomicAdd(&p.part_cursor[b], c);  // ONE global atomic per non-empty bin per tile claims the run
          s_gdelta[b] = g - run;
          uint32_t imax = 0xFFFFFFFFu;
          if (p.part_lim) {
            const uint32_t lim = p.part_lim[b];
            if (g + c > lim) { imax = (g < lim) ? run + (lim - g) : run; atomicOr(p.overflow, 1ull); }
          }
          s_imax[b] = imax;
        }
        run += c;
      }
    }
    __syncthreads();
    const uint32_t total = s_start[trash];  // rows that go somewhere
    // ---- place the rows at their sorted positions (AoS)
#pragma unroll
    for (int k = 0; k < R; k++) {
      const uint32_t b = pid[k];
      if (b == (uint32_t)trash) continue;
      const int row = k * SA_THREADS + tid;
      const uint32_t sp = s_start[b] + s_cnt[warp * n_bins + b] + rank[k];
      s_spid[sp] = (uint16_t)b;
      if constexpr (NC == 2) {
        *reinterpret_cast<ulonglong2 *>(s_sorted + (size_t)sp * 2) = make_ulonglong2(in_s[row], in_s[T + row]);
      } else {
#pragma unroll
        for (int c = 0; c < NC; c++) s_sorted[(size_t)sp * NC + c] = in_s[c * T + row];
      }
    }
    __syncthreads();
    // the input stage is free again: fetch the tile this CTA handles two iterations from now
    if (tid == 0) {
      const int64_t nt = tile + 2 * (int64_t)gridDim.x;
      if (nt < n_tiles && tile_is_tma(nt)) issue(nt, stage);
    }
    // ---- stream the sorted tile out: consecutive threads write consecutive slab rows (full sectors inside a run)
    if constexpr (NC % 2 == 0) {
      constexpr int V = NC / 2;  // 16-byte pieces per row
      for (uint32_t i = tid; i < total * V; i += SA_THREADS) {
        const uint32_t r = i / V, piece = i % V;
        const uint32_t bin = s_spid[r];
        if (r >= s_imax[bin]) continue;  // slab full (the batch is re-run on the exact path)
        const uint64_t dst = (uint64_t)(uint32_t)(s_gdelta[bin] + r);
        tqd::st_stream_u64x2(q.out + dst * NC + piece * 2, *reinterpret_cast<const ulonglong2 *>(s_sorted + (size_t)r * NC + piece * 2));
      }
    } else {
      for (uint32_t i = tid; i < total * NC; i += SA_THREADS) {
        const uint32_t r = i / NC, w = i % NC;
        const uint32_t bin = s_spid[r];
        if (r >= s_imax[bin]) continue;
        const uint64_t dst = (uint64_t)(uint32_t)(s_gdelta[bin] + r);
        tqd::st_stream_u64(q.out + dst * NC + w, s_sorted[(size_t)r * NC + w]);
      }
    }
    // (the next iteration's barriers order these reads before anything they depend on is rewritten)
  }
}

typedef void (*ScatterAosKernel)(const ScatterAosParams);
template <int T>
static ScatterAosKernel scatter_aos_kernel_t(int nc) {
  switch (nc) {
    case 1: return k_scatter_aos<1, T>;
    case 2: return k_scatter_aos<2, T>;
    case 3: return k_scatter_aos<3, T>;
    case 4: return k_scatter_aos<4, T>;
  }
  return nullptr;
}
// tile size: the largest whose shared-memory image fits an SM
static int scatter_aos_tile(int nc, int bins) {
  if (sa_smem_bytes<4096>(nc, bins) <= 220 * 1024 && g_scatter_tile >= 4096) return 4096;
  return 2048;
}
static int32_t launch_scatter_aos(const ScatterAosParams &q, int nc, cudaStream_t s) {
  const int bins = scatter_bins(q.sp) + 1;
  const int T = scatter_aos_tile(nc, bins);
  ScatterAosKernel k = T == 4096 ? scatter_aos_kernel_t<4096>(nc) : scatter_aos_kernel_t<2048>(nc);
  const int smem = T == 4096 ? sa_smem_bytes<4096>(nc, bins) : sa_smem_bytes<2048>(nc, bins);
  if (!k || smem > 227 * 1024) { set_error("internal: AoS scatter of %d columns into %d bins does not fit shared memory", nc, bins); return TQ_ERR_INVALID_ARG; }
  static int attr[2][5] = {};
  if (attr[T == 4096][nc] < smem) {
    TQ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr[T == 4096][nc] = smem;
  }
  const int64_t tiles = (q.sp.n + T - 1) / T;
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int64_t cap = (int64_t)rt().sm_count * per_sm;
  k<<<(int)(tiles < cap ? tiles : cap), SA_THREADS, smem, s>>>(q);
  count_launch();
  return check_launch("k_scatter_aos");
}

// ------------------------------------------------------------------------------------------------ positional probe
// out_base[q] = first output slot of partition q (a multiple of 32), out_base[n_parts] = the span S of the result.
__global__ void k_part_bases(const uint32_t *lo, const uint32_t *hi, const uint32_t *lim, int n_parts, uint32_t *out_base, unsigned long long *span) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t run = 0;
  for (int q = 0; q < n_parts; q++) {
    uint32_t h = hi[q];
    if (lim && h > lim[q]) h = lim[q];
    out_base[q] = run;
    run += ((h - lo[q]) + 31u) & ~31u;
  }
  out_base[n_parts] = run;
  *span = run;
}

static constexpr int PP_STAGES = 4;
static constexpr int PP_TILE = 1024;
static constexpr int PP_CONSUMER_WARPS = 8;
static constexpr int PP_THREADS = (PP_CONSUMER_WARPS + 1) * 32;

struct ProbePosParams {
  const uint64_t *slab;              // AoS probe rows (NP words each), partition q = rows [lo[q], min(hi[q], lim[q]))
  const uint32_t *lo, *hi, *lim;
  const uint32_t *out_base;          // k_part_bases
  uint64_t *out_probe[4], *out_build[4];
  int build_word[4];                 // word of build column c inside a table entry
  uint32_t *valid;                   // one bit per output slot
  unsigned long long *cursor;        // [0] += matched rows
  int key_col, split;
};

template <int NP, int NB>
__global__ void __launch_bounds__(PP_THREADS, 3) k_probe_pos(const ProbePosParams p, const JoinTable t) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  uint64_t *s_tile = reinterpret_cast<uint64_t *>(s_raw);  // [PP_STAGES][PP_TILE][NP]
  __shared__ __align__(8) uint64_t s_full[PP_STAGES], s_empty[PP_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t part = blockIdx.x / p.split, sub = blockIdx.x % p.split;
  const int64_t p_lo = p.lo[part];
  int64_t p_hi = p.hi[part];
  if (p.lim && p_hi > (int64_t)p.lim[part]) p_hi = p.lim[part];
  const int64_t p_tiles = (p_hi - p_lo + PP_TILE - 1) / PP_TILE;
  const int64_t t_lo = p_tiles * sub / p.split, t_hi = p_tiles * (sub + 1) / p.split;
  if (t_lo >= t_hi) return;
  if (tid == 0) {
    for (int s = 0; s < PP_STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], PP_CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == PP_CONSUMER_WARPS) {
    // ---- producer: keeps PP_STAGES slab tiles in flight (TMA bulk copies; bytes rounded up to 16: the slab allocation is padded)
    if (lane == 0) {
      int it = 0;
      for (int64_t tile = t_lo; tile < t_hi; tile++, it++) {
        const int s = it % PP_STAGES;
        if (it >= PP_STAGES) mbar_spin(&s_empty[s], (uint32_t)(((it / PP_STAGES) - 1) & 1));
        const int64_t r0 = p_lo + tile * PP_TILE;
        const int64_t rows = (p_hi - r0) < PP_TILE ? (p_hi - r0) : PP_TILE;
        const uint32_t bytes = (uint32_t)((rows * NP * 8 + 15) & ~15ll);
        mbar_expect_tx(&s_full[s], bytes);
        tma_load_1d(s_tile + (size_t)s * PP_TILE * NP, p.slab + r0 * NP, bytes, &s_full[s]);
      }
    }
    return;
  }
  // ---- consumers: warp w owns rows [w * 128, w * 128 + 128) of every tile
  constexpr int R = PP_TILE / (PP_CONSUMER_WARPS * 32);
  const uint64_t ebase = (uint64_t)part * (t.mask + 1);
  const uint64_t *tbl = t.words + (ebase << t.shift);
  const uint32_t mask = (uint32_t)t.mask;
  const int shift = t.shift;
  const int kc = p.key_col;
  const uint64_t obase = p.out_base[part];
  unsigned matched = 0;
  int it = 0;
  for (int64_t tile = t_lo; tile < t_hi; tile++, it++) {
    const int s = it % PP_STAGES;
    mbar_spin(&s_full[s], (uint32_t)((it / PP_STAGES) & 1));
    const int64_t in_part = tile * PP_TILE + warp * (32 * R) + lane;  // row index inside the partition of this lane's first row
    const uint64_t *src = s_tile + ((size_t)s * PP_TILE + warp * (32 * R) + lane) * NP;
    uint64_t v[R][NP];
#pragma unroll
    for (int k = 0; k < R; k++) {
      if constexpr (NP == 2) {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(src + (size_t)k * 32 * NP);
        v[k][0] = x.x;
        v[k][1] = x.y;
      } else {
#pragma unroll
        for (int c = 0; c < NP; c++) v[k][c] = src[(size_t)k * 32 * NP + c];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);  // the stage can be refilled while this warp probes
    uint64_t key[R];
    uint32_t loc[R];
    bool inb[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      key[k] = v[k][0];
#pragma unroll
      for (int c = 1; c < NP; c++) if (c == kc) key[k] = v[k][c];
      inb[k] = (p_lo + in_part + k * 32) < p_hi;
      loc[k] = home_loc(tqd::mix64(key[k]), mask, shift);
    }
    ulonglong2 ent[R];
    bool hit[R];
    if (shift == 1) {
      EntryPair pr[R];
#pragma unroll
      for (int k = 0; k < R; k++) {  // R independent sector loads in flight
        pr[k].a = make_ulonglong2(EMPTY_KEY, 0);
        pr[k].b = pr[k].a;
        if (inb[k] && key[k] != EMPTY_KEY) pr[k] = ld_pair(tbl, loc[k], false);
      }
#pragma unroll
      for (int k = 0; k < R; k++) {
        hit[k] = false;
        ent[k] = pr[k].a;
        if (inb[k] && key[k] != EMPTY_KEY) {
          for (;;) {
            if (pr[k].a.x == key[k]) { ent[k] = pr[k].a; hit[k] = true; break; }
            if (pr[k].a.x == EMPTY_KEY) break;
            if (pr[k].b.x == key[k]) { ent[k] = pr[k].b; loc[k] += 1; hit[k] = true; break; }
            if (pr[k].b.x == EMPTY_KEY) break;
            loc[k] = (loc[k] + 2) & mask;
            pr[k] = ld_pair(tbl, loc[k], false);
          }
        } else if (inb[k] && t.sent_cnt) {  // a probe key equal to the empty marker: its row is the side entry
          loc[k] = (uint32_t)(t.sent_off - ebase);
          ent[k] = ld_entry(t.words, t.sent_off, 1);
          hit[k] = true;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) {
        ent[k] = make_ulonglong2(EMPTY_KEY, 0);
        if (inb[k] && key[k] != EMPTY_KEY) ent[k] = ld_entry(tbl, loc[k], shift);
      }
#pragma unroll
      for (int k = 0; k < R; k++) {
        hit[k] = false;
        if (inb[k] && key[k] != EMPTY_KEY) {
          while (ent[k].x != key[k] && ent[k].x != EMPTY_KEY) {
            loc[k] = (loc[k] + 1) & mask;
            ent[k] = ld_entry(tbl, loc[k], shift);
          }
          hit[k] = ent[k].x == key[k];
        } else if (inb[k] && t.sent_cnt) {
          loc[k] = (uint32_t)(t.sent_off - ebase);
          ent[k] = ld_entry(t.words, t.sent_off, shift);
          hit[k] = true;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) {
      const uint64_t o = obase + (uint64_t)(in_part + k * 32);  // this row's output slot: its position in the partition order
      const unsigned bal = __ballot_sync(0xffffffffu, hit[k]);
      // slots past the partition's padded end belong to the next partition: only groups that start inside it are written
      const bool group_live = (p_lo + in_part - lane + k * 32) < p_hi;
      if (lane == 0 && group_live) {
        p.valid[(o - lane) >> 5] = bal;
        matched += __popc(bal);
      }
      if (hit[k]) {
#pragma unroll
        for (int c = 0; c < NP; c++) tqd::st_stream_u64(p.out_probe[c] + o, v[k][c]);
#pragma unroll
        for (int c = 0; c < NB; c++) {
          uint64_t x;
          if (p.build_word[c] == 0) x = ent[k].x;
          else if (p.build_word[c] == 1) x = ent[k].y;
          else x = t.words[((ebase + loc[k]) << shift) + p.build_word[c]];  // words 2..3 of a 32-byte entry: same sector
          tqd::st_stream_u64(p.out_build[c] + o, x);
        }
      }
    }
  }
  if (lane == 0 && matched) atomicAdd(p.cursor, (unsigned long long)matched);
}

typedef void (*ProbePosKernel)(const ProbePosParams, const JoinTable);
template <int NP>
static ProbePosKernel probe_pos_nb(int nb) {
  switch (nb) {
    case 1: return k_probe_pos<NP, 1>;
    case 2: return k_probe_pos<NP, 2>;
    case 3: return k_probe_pos<NP, 3>;
    case 4: return k_probe_pos<NP, 4>;
  }
  return nullptr;
}
static ProbePosKernel probe_pos_kernel(int np, int nb) {
  switch (np) {
    case 1: return probe_pos_nb<1>(nb);
    case 2: return probe_pos_nb<2>(nb);
    case 3: return probe_pos_nb<3>(nb);
    case 4: return probe_pos_nb<4>(nb);
  }
  return nullptr;
}

// ---- hole filling: the k-th empty slot below M takes the k-th real row at or above M (M = rows of the result)
__global__ void __launch_bounds__(256) k_hole_popc(const uint32_t *valid, int64_t n_words, uint32_t *cnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) cnt[w] = __popc(valid[w]);
}
__device__ __forceinline__ uint32_t valid_rank(const uint32_t *valid, const uint32_t *vpre, uint64_t pos, int64_t n_words) {
  const uint64_t w = pos >> 5;
  if ((int64_t)w >= n_words) return vpre[n_words];  // vpre has n_words + 1 entries (the last = total)
  return vpre[w] + __popc(valid[w] & ((1u << (pos & 31)) - 1u));
}
__global__ void __launch_bounds__(256) k_hole_lists(const uint32_t *valid, const uint32_t *vpre, int64_t n_words, uint64_t M, uint32_t *hole_pos, uint32_t *tail_src) {
  const uint32_t below = valid_rank(valid, vpre, M, n_words);  // real rows in [0, M)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    const uint32_t bits = valid[w];
    const uint64_t p0 = (uint64_t)w << 5;
    if (p0 + 32 <= M && bits == 0xFFFFFFFFu) continue;
    if (p0 >= M && bits == 0) continue;
    const uint32_t pre = vpre[w];
    for (int b = 0; b < 32; b++) {
      const uint64_t pos = p0 + b;
      const bool v = (bits >> b) & 1u;
      const uint32_t vr = pre + __popc(bits & ((1u << b) - 1u));
      if (pos < M && !v) hole_pos[pos - vr] = (uint32_t)pos;             // holes before pos = pos - (real rows before pos)
      else if (pos >= M && v) tail_src[vr - below] = (uint32_t)pos;
    }
  }
}
struct HoleMoveParams {
  int n_cols;
  uint64_t *col[8];
  const uint32_t *hole_pos, *tail_src;
  const uint32_t *valid, *vpre;
  int64_t n_words;
  uint64_t M;
};
__global__ void __launch_bounds__(256) k_hole_move(const HoleMoveParams h) {
  const uint64_t H = h.M - valid_rank(h.valid, h.vpre, h.M, h.n_words);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < H; k += stride) {
    const uint32_t d = h.hole_pos[k], s = h.tail_src[k];
    for (int c = 0; c < h.n_cols; c++) h.col[c][d] = h.col[c][s];
  }
}

// ------------------------------------------------------------------------------------------------ partition-local build
static constexpr int BP_THREADS = 1024;
struct BuildPartParams {
  const uint64_t *slab;              // AoS build rows (NB words each)
  const uint32_t *lo, *hi;
  uint64_t *words;                   // the table
  uint64_t cap;                      // entries per partition table (power of two)
  int shift, n_parts, key_col;
  int word_of_col[4];
  unsigned *flags;                   // |= 1: duplicate key, |= 2: the empty-marker key appeared  -> the caller rebuilds on the general path
};
template <int NB>
__global__ void __launch_bounds__(BP_THREADS, 1) k_build_part(const BuildPartParams b) {
  const int tid = threadIdx.x;
  const uint64_t mask = b.cap - 1;
  for (int part = blockIdx.x; part < b.n_parts; part += gridDim.x) {
    uint64_t *tbl = b.words + (((uint64_t)part * b.cap) << b.shift);
    // ---- init: (EMPTY_KEY, 0[, 0, 0]) entries, 16-byte stores; the table partition stays in L2 for the inserts below
    const uint64_t n_vec = (b.cap << b.shift) >> 1;
    const int per_entry = 1 << (b.shift - 1);  // 16-byte pieces per entry
    for (uint64_t i = tid; i < n_vec; i += BP_THREADS) {
      const bool first = (i & (uint64_t)(per_entry - 1)) == 0;
      *reinterpret_cast<ulonglong2 *>(tbl + i * 2) = make_ulonglong2(first ? EMPTY_KEY : 0ull, 0ull);
    }
    __syncthreads();
    const int64_t lo = b.lo[part], hi = b.hi[part];
    for (int64_t r = lo + tid; r < hi; r += BP_THREADS) {
      uint64_t w[NB];
      if constexpr (NB == 2) {
        const ulonglong2 x = tqd::ld_stream_u64x2(b.slab + r * 2);
        w[0] = x.x;
        w[1] = x.y;
      } else {
#pragma unroll
        for (int c = 0; c < NB; c++) w[c] = tqd::ld_stream_u64(b.slab + r * NB + c);
      }
      uint64_t key = w[0];
#pragma unroll
      for (int c = 1; c < NB; c++) if (c == b.key_col) key = w[c];
      if (key == EMPTY_KEY) { atomicOr(b.flags, 2u); continue; }
      const uint64_t h = tqd::mix64(key);
      uint64_t loc = (b.shift == 1) ? ((h & mask) & ~1ull) : (h & mask);
      for (;;) {
        uint64_t *ent = tbl + (loc << b.shift);
        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(ent), (unsigned long long)EMPTY_KEY, (unsigned long long)key);
        if (prev == EMPTY_KEY) {
#pragma unroll
          for (int c = 0; c < NB; c++) if (c != b.key_col) ent[b.word_of_col[c]] = w[c];
          break;
        }
        if (prev == key) { atomicOr(b.flags, 1u); break; }
        loc = (loc + 1) & mask;
      }
    }
    // (no barrier needed before the next partition: it lives elsewhere)
  }
}
typedef void (*BuildPartKernel)(const BuildPartParams);
static BuildPartKernel build_part_kernel(int nb) {
  switch (nb) {
    case 1: return k_build_part<1>;
    case 2: return k_build_part<2>;
    case 3: return k_build_part<3>;
    case 4: return k_build_part<4>;
  }
  return nullptr;
}

}  // namespace tq
