/*
 * cpu_ref.c — multi-threaded CPU restatement of the reference's goroutine DESIGN, used only as the
 * cpu_baseline / `--impl reference` arm of bench.py (TEST/BENCH INFRASTRUCTURE, never the product).
 *
 * The Go reference cannot be built here (no Go toolchain; join/agg hot functions are stubs), so this
 * restates its structure: executor/join.go (serial build, N probe workers over 1024-row chunks with
 * private result chunks), executor/hash_table.go (hash -> chained entries, insertion-order Get,
 * key re-verification), executor/aggregate.go (P partial workers with private maps, hash shuffle to
 * F final workers), expression/builtin_compare_vec.go + builtin_arithmetic_vec.go (1024-row loops).
 * Results are validated against oracle.c in tests/test_oracle_golden.py.
 */
#include "oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CHUNK 1024 /* DefMaxChunkSize sessionctx/variable/tidb_vars.go:241 */

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* FNV-1 64 of varintFlag || 8 raw bytes (hash_table.go:64, codec.go:258-274) */
static inline uint64_t fnv_key(int64_t k) {
  uint64_t h = 14695981039346656037ULL;
  h = (h * 1099511628211ULL) ^ 8u;
  uint64_t r = (uint64_t)k;
  for (int b = 0; b < 8; b++) h = (h * 1099511628211ULL) ^ (uint8_t)(r >> (8 * b));
  return h;
}

/* ---------------------------------------------------------------- join */
typedef struct { int64_t row; int64_t next; } jentry;
typedef struct {
  uint64_t *bkt_hash; int64_t *bkt_head; int64_t n_bkt; /* map[uint64]entryAddr */
  jentry *entries;
  const int64_t *bk, *bv, *pk, *pv; int64_t n_probe;
  int workers;
} jshared;
typedef struct { jshared *s; int id; int64_t rows; uint64_t checksum; } jworker;

static void *join_worker(void *arg) {
  jworker *w = (jworker *)arg; jshared *s = w->s;
  /* private result chunk: 4 columns x 1024 rows (join.go:255-266 getNewJoinResult) */
  int64_t *res = (int64_t *)malloc(sizeof(int64_t) * 4 * CHUNK);
  int64_t nres = 0, total = 0; uint64_t cks = 0;
  int64_t matched[64]; int64_t *mbuf = matched; int64_t mcap = 64;
  int64_t n_chunks = (s->n_probe + CHUNK - 1) / CHUNK;
  for (int64_t c = w->id; c < n_chunks; c += s->workers) {       /* outer chunks dealt to workers (join.go:194-221) */
    int64_t lo = c * CHUNK, hi = lo + CHUNK; if (hi > s->n_probe) hi = s->n_probe;
    uint64_t hv[CHUNK];
    for (int64_t i = lo; i < hi; i++) hv[i - lo] = fnv_key(s->pk[i]);  /* HashChunkSelected over the chunk (join.go:335-341) */
    for (int64_t i = lo; i < hi; i++) {                           /* per-row loop (join.go:343-360) */
      uint64_t h = hv[i - lo];
      int64_t b = (int64_t)((h * 0x9E3779B97F4A7C15ULL) >> 17) & (s->n_bkt - 1);
      while (s->bkt_head[b] != -1 && s->bkt_hash[b] != h) b = (b + 1) & (s->n_bkt - 1);
      int64_t nm = 0;
      for (int64_t e = s->bkt_head[b]; e != -1; e = s->entries[e].next) {  /* rowHashMap.Get (hash_table.go:259-272) */
        if (nm == mcap) { mcap *= 2; int64_t *nb = (int64_t *)malloc(8 * (size_t)mcap); memcpy(nb, mbuf, 8 * (size_t)nm); if (mbuf != matched) free(mbuf); mbuf = nb; }
        mbuf[nm++] = s->entries[e].row;
      }
      for (int64_t j = nm - 1; j >= 0; j--) {                     /* insertion order; matchJoinKey (hash_table.go:137-141) */
        int64_t br = mbuf[j];
        if (s->bk[br] != s->pk[i]) continue;
        res[0 * CHUNK + nres] = s->bk[br]; res[1 * CHUNK + nres] = s->bv[br];   /* makeJoinRowToChunk: inner ++ outer (outerIsRight) */
        res[2 * CHUNK + nres] = s->pk[i];  res[3 * CHUNK + nres] = s->pv[i];
        if (++nres == CHUNK) {                                     /* chk.IsFull -> joinResultCh (join.go:353-359) */
          for (int64_t r = 0; r < CHUNK; r++) cks += (uint64_t)res[1 * CHUNK + r] ^ (uint64_t)res[3 * CHUNK + r];
          total += nres; nres = 0;
        }
      }
    }
  }
  for (int64_t r = 0; r < nres; r++) cks += (uint64_t)res[1 * CHUNK + r] ^ (uint64_t)res[3 * CHUNK + r];
  total += nres;
  if (mbuf != matched) free(mbuf);
  free(res);
  w->rows = total; w->checksum = cks;
  return NULL;
}

int64_t orc_mt_join_bench(int64_t n_build, const int64_t *bk, const int64_t *bv, int64_t n_probe, const int64_t *pk,
                          const int64_t *pv, int workers, double *build_seconds, double *probe_seconds, uint64_t *checksum) {
  jshared s; memset(&s, 0, sizeof(s));
  s.bk = bk; s.bv = bv; s.pk = pk; s.pv = pv; s.n_probe = n_probe; s.workers = workers < 1 ? 1 : workers;
  double t0 = now_s();
  /* single-threaded build: hashRowContainer.PutChunk per 1024-row chunk (hash_table.go:146-169, "not thread-safe") */
  s.n_bkt = 1024; while (s.n_bkt < 2 * n_build) s.n_bkt <<= 1;
  s.bkt_hash = (uint64_t *)malloc(8 * (size_t)s.n_bkt); s.bkt_head = (int64_t *)malloc(8 * (size_t)s.n_bkt);
  for (int64_t i = 0; i < s.n_bkt; i++) s.bkt_head[i] = -1;
  s.entries = (jentry *)malloc(sizeof(jentry) * (size_t)(n_build ? n_build : 1));
  for (int64_t i = 0; i < n_build; i++) {
    uint64_t h = fnv_key(bk[i]);
    int64_t b = (int64_t)((h * 0x9E3779B97F4A7C15ULL) >> 17) & (s.n_bkt - 1);
    while (s.bkt_head[b] != -1 && s.bkt_hash[b] != h) b = (b + 1) & (s.n_bkt - 1);
    s.entries[i].row = i; s.entries[i].next = s.bkt_head[b];       /* rowHashMap.Put (hash_table.go:247-256) */
    s.bkt_hash[b] = h; s.bkt_head[b] = i;
  }
  double t1 = now_s();
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)s.workers);
  jworker *ws = (jworker *)calloc((size_t)s.workers, sizeof(jworker));
  for (int w = 0; w < s.workers; w++) { ws[w].s = &s; ws[w].id = w; pthread_create(&th[w], NULL, join_worker, &ws[w]); }
  int64_t total = 0; uint64_t cks = 0;
  for (int w = 0; w < s.workers; w++) { pthread_join(th[w], NULL); total += ws[w].rows; cks += ws[w].checksum; }
  double t2 = now_s();
  if (build_seconds) *build_seconds = t1 - t0;
  if (probe_seconds) *probe_seconds = t2 - t1;
  if (checksum) *checksum = cks;
  free(th); free(ws); free(s.bkt_hash); free(s.bkt_head); free(s.entries);
  return total;
}

/* ---------------------------------------------------------------- agg */
typedef struct { int64_t *keys; double *sums; int64_t *cnts; uint8_t *used; int64_t cap, n; } gmap;
static void gmap_init(gmap *m, int64_t cap) {
  m->cap = 1024; while (m->cap < cap) m->cap <<= 1; m->n = 0;
  m->keys = (int64_t *)malloc(8 * (size_t)m->cap); m->sums = (double *)malloc(8 * (size_t)m->cap);
  m->cnts = (int64_t *)malloc(8 * (size_t)m->cap); m->used = (uint8_t *)calloc((size_t)m->cap, 1);
}
static void gmap_free(gmap *m) { free(m->keys); free(m->sums); free(m->cnts); free(m->used); }
static inline int64_t gmap_slot(gmap *m, int64_t k);
static void gmap_grow(gmap *m) {
  gmap o = *m; gmap_init(m, o.cap * 2);
  for (int64_t i = 0; i < o.cap; i++) if (o.used[i]) { int64_t s = gmap_slot(m, o.keys[i]); m->used[s] = 1; m->keys[s] = o.keys[i]; m->sums[s] = o.sums[i]; m->cnts[s] = o.cnts[i]; m->n++; }
  gmap_free(&o);
}
static inline int64_t gmap_slot(gmap *m, int64_t k) {
  int64_t s = (int64_t)((fnv_key(k) * 0x9E3779B97F4A7C15ULL) >> 17) & (m->cap - 1);
  while (m->used[s] && m->keys[s] != k) s = (s + 1) & (m->cap - 1);
  return s;
}
typedef struct { const int64_t *k; const double *x; int64_t n; int P, F; gmap *partial; gmap *fin; } ashared;
typedef struct { ashared *s; int id; } aworker;
static void *agg_partial(void *arg) {
  aworker *w = (aworker *)arg; ashared *s = w->s; gmap *m = &s->partial[w->id];
  int64_t n_chunks = (s->n + CHUNK - 1) / CHUNK;
  for (int64_t c = w->id; c < n_chunks; c += s->P) {              /* fetchChildData -> partialInputChs (aggregate.go:487-522) */
    int64_t lo = c * CHUNK, hi = lo + CHUNK; if (hi > s->n) hi = s->n;
    for (int64_t i = lo; i < hi; i++) {                           /* updatePartialResult (aggregate.go:332-350) */
      if ((m->n + 1) * 2 > m->cap) gmap_grow(m);
      int64_t sl = gmap_slot(m, s->k[i]);
      if (!m->used[sl]) { m->used[sl] = 1; m->keys[sl] = s->k[i]; m->sums[sl] = s->x[i]; m->cnts[sl] = 1; m->n++; }
      else { m->sums[sl] += s->x[i]; m->cnts[sl]++; }
    }
  }
  return NULL;
}
static void *agg_final(void *arg) {
  aworker *w = (aworker *)arg; ashared *s = w->s; gmap *m = &s->fin[w->id];
  for (int p = 0; p < s->P; p++) {                                /* shuffleIntermData / consumeIntermData (aggregate.go:352-356,424-427) */
    gmap *src = &s->partial[p];
    for (int64_t i = 0; i < src->cap; i++) {
      if (!src->used[i]) continue;
      if ((int)(fnv_key(src->keys[i]) % (uint64_t)s->F) != w->id) continue;
      if ((m->n + 1) * 2 > m->cap) gmap_grow(m);
      int64_t sl = gmap_slot(m, src->keys[i]);
      if (!m->used[sl]) { m->used[sl] = 1; m->keys[sl] = src->keys[i]; m->sums[sl] = src->sums[i]; m->cnts[sl] = src->cnts[i]; m->n++; }
      else { m->sums[sl] += src->sums[i]; m->cnts[sl] += src->cnts[i]; }   /* MergePartialResult */
    }
  }
  return NULL;
}
int64_t orc_mt_agg_bench(int64_t n, const int64_t *k, const double *x, int P, int F, double *seconds, double *sum_of_sums, int64_t *sum_of_counts) {
  if (P < 1) P = 1; if (F < 1) F = 1;
  ashared s; s.k = k; s.x = x; s.n = n; s.P = P; s.F = F;
  s.partial = (gmap *)malloc(sizeof(gmap) * (size_t)P); s.fin = (gmap *)malloc(sizeof(gmap) * (size_t)F);
  double t0 = now_s();
  for (int p = 0; p < P; p++) gmap_init(&s.partial[p], 1024);
  for (int f = 0; f < F; f++) gmap_init(&s.fin[f], 1024);
  int T = P > F ? P : F;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)T); aworker *ws = (aworker *)malloc(sizeof(aworker) * (size_t)T);
  for (int p = 0; p < P; p++) { ws[p].s = &s; ws[p].id = p; pthread_create(&th[p], NULL, agg_partial, &ws[p]); }
  for (int p = 0; p < P; p++) pthread_join(th[p], NULL);
  for (int f = 0; f < F; f++) { ws[f].s = &s; ws[f].id = f; pthread_create(&th[f], NULL, agg_final, &ws[f]); }
  for (int f = 0; f < F; f++) pthread_join(th[f], NULL);
  double t1 = now_s();
  int64_t groups = 0, cnts = 0; double sums = 0;
  for (int f = 0; f < F; f++) { groups += s.fin[f].n; for (int64_t i = 0; i < s.fin[f].cap; i++) if (s.fin[f].used[i]) { sums += s.fin[f].sums[i]; cnts += s.fin[f].cnts[i]; } }
  if (seconds) *seconds = t1 - t0; if (sum_of_sums) *sum_of_sums = sums; if (sum_of_counts) *sum_of_counts = cnts;
  for (int p = 0; p < P; p++) gmap_free(&s.partial[p]); for (int f = 0; f < F; f++) gmap_free(&s.fin[f]);
  free(s.partial); free(s.fin); free(th); free(ws);
  return groups;
}

/* ---------------------------------------------------------------- LT + Plus */
typedef struct { const int64_t *a, *b; int64_t *lt, *plus; int64_t n; int workers, id; int64_t overflow; } eworker;
static void *expr_worker(void *arg) {
  eworker *w = (eworker *)arg;
  int64_t n_chunks = (w->n + CHUNK - 1) / CHUNK;
  for (int64_t c = w->id; c < n_chunks; c += w->workers) {        /* projection workers over chunks (projection.go:209-256) */
    int64_t lo = c * CHUNK, hi = lo + CHUNK; if (hi > w->n) hi = w->n;
    const int64_t *a = w->a, *b = w->b; int64_t *r = w->lt, *p = w->plus;
    for (int64_t i = lo; i < hi; i++) r[i] = a[i] < b[i] ? -1 : (a[i] == b[i] ? 0 : 1);  /* VecCompareII types/compare.go:58-69 */
    for (int64_t i = lo; i < hi; i++) r[i] = r[i] < 0 ? 1 : 0;                            /* vecResOfLT builtin_compare_vec.go:214-223 */
    for (int64_t i = lo; i < hi; i++) {                                                   /* plusSS builtin_arithmetic_vec.go:481-495 */
      int64_t lh = a[i], rh = b[i];
      if ((lh > 0 && rh > INT64_MAX - lh) || (lh < 0 && rh < INT64_MIN - lh)) w->overflow++;
      p[i] = lh + rh;
    }
  }
  return NULL;
}
int64_t orc_mt_lt_plus_bench(int64_t n, const int64_t *a, const int64_t *b, int64_t *lt_out, int64_t *plus_out, int workers, double *seconds) {
  if (workers < 1) workers = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)workers); eworker *ws = (eworker *)calloc((size_t)workers, sizeof(eworker));
  double t0 = now_s();
  for (int w = 0; w < workers; w++) { ws[w].a = a; ws[w].b = b; ws[w].lt = lt_out; ws[w].plus = plus_out; ws[w].n = n; ws[w].workers = workers; ws[w].id = w; pthread_create(&th[w], NULL, expr_worker, &ws[w]); }
  int64_t ovf = 0;
  for (int w = 0; w < workers; w++) { pthread_join(th[w], NULL); ovf += ws[w].overflow; }
  if (seconds) *seconds = now_s() - t0;
  free(th); free(ws);
  return ovf;
}
