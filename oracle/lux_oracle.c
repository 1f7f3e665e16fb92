/*
 * lux_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A Legion-free restatement, in plain C + OpenMP, of the semantics of the LuxGraph/Lux hot path
 * (reference checkout /root/reference @ 6263711).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (lux_b200/csrc) never links or calls it.
 *
 * PARITY PIN STATUS: PINNED BY REFERENCE EXECUTION.  The reference ships no golden vectors, no tests and no CPU
 * compute path and its driver needs Legion, but its task bodies and CUDA kernels compile
 * unmodified from /root/reference behind a small Legion shim (oracle/ref_replay/ -> oracle/_ref/libref_*.so) and
 * its converter compiles as is (oracle/build_ref.py -> oracle/_ref/converter).  Pins:
 *   (1) tests/golden/ref_replay_golden.npz — outputs of the reference's OWN pagerank / components / sssp kernels
 *       replayed on a B200 (scripts/make_ref_golden.py): CC / SSSP labels bit-exact, iteration counts and active
 *       counts equal (except the reference's double count after a frontier promotion, defect B5), PageRank within
 *       the reference's own float-atomicAdd noise (1e-6 on low-degree graphs, 5e-5 on a 10^4 hub);
 *   (2) tests/golden/hand5.lux.hex — bytes written by the reference's tools/converter.cc;
 *   (3) hand-derived answers and an independent numpy restatement (tests/test_oracle.py).
 * col_filter is NOT pinned by execution: the reference kernel is racy and mis-indexed; the oracle
 * restates the intended math and is pinned by (3) and MATHEMATICALLY by tests/test_oracle.py::
 * test_colfilter_step_is_minus_gamma_times_the_gradient_...: lo_cf_iter's step equals -GAMMA times the
 * finite-difference gradient of the regularised squared error the reference's constants define (col_filter/app.h:26-28).
 *
 * Every function cites the reference file:line it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint32_t V_ID;   /* pagerank/app.h:21, components/app.h:21 */
typedef uint64_t E_ID;   /* pagerank/app.h:22 */

#define LO_ALPHA 0.15f          /* pagerank/app.h:24 */
#define LO_CF_K 20              /* col_filter/app.h:28 */
#define LO_CF_LAMBDA 0.001f     /* col_filter/app.h:26 */
#define LO_CF_GAMMA 0.00000035f /* col_filter/app.h:27 */
#define LO_SPARSE_THRESHOLD 16  /* components/app.h:19 */
#define LO_DENSE_BITMAP 0x1234567u /* core/graph.h:102 */
#define LO_SPARSE_QUEUE 0x7654321u /* core/graph.h:103 */

enum { LO_APP_CC = 1, LO_APP_SSSP = 2 };

void lo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int lo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Deterministic synthetic inputs (ours — the reference has no generator).
 * Counter-based: edge i depends only on (seed, i) so CPU and GPU build identical graphs.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
uint64_t lo_splitmix64(uint64_t x) { return splitmix64(x); }

/* RMAT quadrant thresholds in 16-bit fixed point: a=.57 b=.19 c=.19 d=.05 (Graph500). */
#define RMAT_T0 37356u /* a            */
#define RMAT_T1 49807u /* a+b          */
#define RMAT_T2 62259u /* a+b+c        */

/* One RMAT edge; rejection of endpoints >= nv (for non power-of-two "X-scale" graphs). */
void lo_rmat_edge(uint64_t seed, uint64_t i, int scale, V_ID nv, V_ID* src_out, V_ID* dst_out) {
  uint64_t h0 = splitmix64(splitmix64(seed) ^ i);
  for (uint64_t attempt = 0;; attempt++) {
    uint64_t ha = splitmix64(h0 + attempt);
    uint32_t s = 0, d = 0;
    uint64_t w = 0;
    for (int lvl = 0; lvl < scale; lvl++) {
      if ((lvl & 3) == 0) w = splitmix64(ha ^ ((uint64_t)(lvl / 4 + 1) * 0xA0761D6478BD642Full));
      uint32_t r = (uint32_t)(w & 0xFFFFu);
      w >>= 16;
      uint32_t sb, db;
      if (r < RMAT_T0) { sb = 0; db = 0; }
      else if (r < RMAT_T1) { sb = 0; db = 1; }
      else if (r < RMAT_T2) { sb = 1; db = 0; }
      else { sb = 1; db = 1; }
      s = (s << 1) | sb;
      d = (d << 1) | db;
    }
    if (s < nv && d < nv) { *src_out = s; *dst_out = d; return; }
  }
}

/* Edge weight as a pure function of (seed, src, dst): uniform int 1..5 (NetFlix-like ratings). */
int32_t lo_edge_weight(uint64_t seed, V_ID src, V_ID dst) {
  uint64_t h = splitmix64(splitmix64(seed ^ 0x5bd1e995u) ^ (((uint64_t)dst << 32) | src));
  return (int32_t)(1 + (h >> 33) % 5);
}

/* Bipartite rating j -> (user, item) with a skewed item distribution; integer-only arithmetic. */
void lo_bipartite_edge(uint64_t seed, uint64_t j, V_ID users, V_ID items, V_ID* user_out, V_ID* item_out) {
  uint64_t h1 = splitmix64(splitmix64(seed) ^ j);
  uint64_t h2 = splitmix64(h1 ^ 0xA0761D6478BD642Full);
  V_ID user = (V_ID)(((h1 >> 32) * (uint64_t)users) >> 32);
  uint64_t a = h2 & 0xFFFFFFFFull, b = h2 >> 32;
  uint64_t m = (a * b) >> 32; /* product of two uniforms: density -ln(x), skewed to 0 */
  V_ID item = (V_ID)((m * (uint64_t)items) >> 32);
  *user_out = user;
  *item_out = users + item;
}

/* LSD radix sort of 64-bit keys (16-bit digits) — canonical CSC order is (dst, src) ascending. */
static void radix_sort_u64(uint64_t* keys, uint64_t n, int key_bits) {
  uint64_t* tmp = (uint64_t*)malloc(n * sizeof(uint64_t));
  uint64_t* cnt = (uint64_t*)malloc(65536 * sizeof(uint64_t));
  uint64_t *a = keys, *b = tmp;
  for (int shift = 0; shift < key_bits; shift += 16) {
    memset(cnt, 0, 65536 * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) cnt[(a[i] >> shift) & 0xFFFF]++;
    uint64_t run = 0;
    for (int k = 0; k < 65536; k++) { uint64_t c = cnt[k]; cnt[k] = run; run += c; }
    for (uint64_t i = 0; i < n; i++) b[cnt[(a[i] >> shift) & 0xFFFF]++] = a[i];
    uint64_t* t = a; a = b; b = t;
  }
  if (a != keys) memcpy(keys, a, n * sizeof(uint64_t));
  free(tmp);
  free(cnt);
}

/* keys (dst<<32|src), sorted -> CSC arrays in the reference's on-disk convention:
 * row_end[v] = END offset of v's in-edge block (tools/converter.cc:100-106, pull_model.inl:99-102). */
static void keys_to_csc(const uint64_t* keys, uint64_t ne, V_ID nv, E_ID* row_end, V_ID* src) {
  uint64_t e = 0;
  for (V_ID v = 0; v < nv; v++) {
    while (e < ne && (V_ID)(keys[e] >> 32) == v) { src[e] = (V_ID)(keys[e] & 0xFFFFFFFFu); e++; }
    row_end[v] = e;
  }
}

int lo_gen_rmat_csc(int scale, V_ID nv, E_ID ne, uint64_t seed, E_ID* row_end, V_ID* src) {
  uint64_t* keys = (uint64_t*)malloc(ne * sizeof(uint64_t));
  if (!keys) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)ne; i++) {
    V_ID s, d;
    lo_rmat_edge(seed, (uint64_t)i, scale, nv, &s, &d);
    keys[i] = ((uint64_t)d << 32) | s;
  }
  radix_sort_u64(keys, ne, 64);
  keys_to_csc(keys, ne, nv, row_end, src);
  free(keys);
  return 0;
}

/* ---- sampled CSC of the RMAT graph: only the destination BLOCKS (2^block_shift consecutive ids) flagged in
 * block_sel are materialised, in ascending block order ("compact" local vertex numbering).  Used at full scale
 * (RMAT-27) by bench.py's CPU baseline / reference arm and its N-GPU parity check, where building the whole CSC on
 * the host would cost minutes: one OpenMP scan of the edge counter per pass, arrays first-touched by the threads
 * that fill them (NUMA-interleaved).
 *   pass 1  lo_rmat_blocks_count: indeg_local[i] (per compact vertex) and, if deg != NULL, the global out-degrees
 *           (pull_scan_task_impl, pull_model.inl:333-343);
 *   pass 2  lo_rmat_blocks_fill : src[] for the compact vertices, canonical order (ascending source inside a vertex). */
static inline int64_t block_local_base(int block_shift, const int64_t* block_base, V_ID d) {
  return block_base[d >> block_shift];
}

int lo_rmat_blocks_count(int scale, V_ID nv, E_ID ne, uint64_t seed, int block_shift, const uint8_t* block_sel,
                         V_ID* deg /* [nv] or NULL */, V_ID* indeg_local /* [n_local] */, uint64_t n_local) {
  const uint64_t n_blocks = (((uint64_t)nv - 1) >> block_shift) + 1;
  int64_t* block_base = (int64_t*)malloc(n_blocks * sizeof(int64_t));
  if (!block_base) return -1;
  int64_t run = 0;
  for (uint64_t b = 0; b < n_blocks; b++) {
    if (block_sel[b]) {
      uint64_t lo = b << block_shift, hi = lo + (1ull << block_shift);
      if (hi > nv) hi = nv;
      block_base[b] = run;
      run += (int64_t)(hi - lo);
    } else block_base[b] = -1;
  }
  if ((uint64_t)run != n_local) { free(block_base); return -2; }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n_local; i++) indeg_local[i] = 0;
  if (deg) {
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < (int64_t)nv; v++) deg[v] = 0;
  }
  const uint64_t mask = (1ull << block_shift) - 1;
  /* Counters of hub vertices are hammered by every thread (vertex 0 of RMAT-27 is an endpoint of 1.3 M edges): each
   * thread combines its increments in a small direct-mapped cache and only touches the shared array on eviction. */
#pragma omp parallel
  {
    enum { CACHE = 8192 };
    struct { V_ID* addr; V_ID cnt; } *cs = calloc(CACHE, sizeof(*cs)), *cd = calloc(CACHE, sizeof(*cd));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)ne; i++) {
      V_ID s, d;
      lo_rmat_edge(seed, (uint64_t)i, scale, nv, &s, &d);
      if (deg) {
        uint32_t slot = (s * 2654435761u) >> 19;
        if (cs[slot].addr != &deg[s]) {
          if (cs[slot].cnt) {
#pragma omp atomic
            *cs[slot].addr += cs[slot].cnt;
          }
          cs[slot].addr = &deg[s];
          cs[slot].cnt = 0;
        }
        cs[slot].cnt++;
      }
      int64_t base = block_local_base(block_shift, block_base, d);
      if (base >= 0) {
        V_ID* a = &indeg_local[base + (int64_t)(d & mask)];
        uint32_t slot = (d * 2654435761u) >> 19;
        if (cd[slot].addr != a) {
          if (cd[slot].cnt) {
#pragma omp atomic
            *cd[slot].addr += cd[slot].cnt;
          }
          cd[slot].addr = a;
          cd[slot].cnt = 0;
        }
        cd[slot].cnt++;
      }
    }
    for (int k = 0; k < CACHE; k++) {
      if (cs[k].cnt) {
#pragma omp atomic
        *cs[k].addr += cs[k].cnt;
      }
      if (cd[k].cnt) {
#pragma omp atomic
        *cd[k].addr += cd[k].cnt;
      }
    }
    free(cs);
    free(cd);
  }
  free(block_base);
  return 0;
}

static int cmp_vid(const void* a, const void* b) {
  V_ID x = *(const V_ID*)a, y = *(const V_ID*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

int lo_rmat_blocks_fill(int scale, V_ID nv, E_ID ne, uint64_t seed, int block_shift, const uint8_t* block_sel,
                        const E_ID* row_end_local /* [n_local] inclusive scan of indeg_local */, uint64_t n_local,
                        V_ID* src /* [row_end_local[n_local-1]] */) {
  const uint64_t n_blocks = (((uint64_t)nv - 1) >> block_shift) + 1;
  int64_t* block_base = (int64_t*)malloc(n_blocks * sizeof(int64_t));
  E_ID* cursor = (E_ID*)malloc((n_local ? n_local : 1) * sizeof(E_ID));
  if (!block_base || !cursor) { free(block_base); free(cursor); return -1; }
  int64_t run = 0;
  for (uint64_t b = 0; b < n_blocks; b++) {
    if (block_sel[b]) {
      uint64_t lo = b << block_shift, hi = lo + (1ull << block_shift);
      if (hi > nv) hi = nv;
      block_base[b] = run;
      run += (int64_t)(hi - lo);
    } else block_base[b] = -1;
  }
  if ((uint64_t)run != n_local) { free(block_base); free(cursor); return -2; }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n_local; i++) cursor[i] = i == 0 ? 0 : row_end_local[i - 1];
  const uint64_t mask = (1ull << block_shift) - 1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)ne; i++) {
    V_ID s, d;
    lo_rmat_edge(seed, (uint64_t)i, scale, nv, &s, &d);
    int64_t base = block_local_base(block_shift, block_base, d);
    if (base >= 0) {
      E_ID pos;
      E_ID* c = &cursor[base + (int64_t)(d & mask)];
#pragma omp atomic capture
      pos = (*c)++;
      src[pos] = s;
    }
  }
  /* canonical order inside each vertex: ascending source (the slots were claimed in thread-arrival order) */
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < (int64_t)n_local; i++) {
    E_ID b = i == 0 ? 0 : row_end_local[i - 1], e = row_end_local[i];
    if (e - b > 1) qsort(src + b, (size_t)(e - b), sizeof(V_ID), cmp_vid);
  }
  free(block_base);
  free(cursor);
  return 0;
}

/* ratings stored in both directions: ne must be 2*ratings. weights filled from lo_edge_weight of the
 * (user,item) pair so both directions carry the same rating. */
int lo_gen_bipartite_csc(V_ID users, V_ID items, E_ID ratings, uint64_t seed, E_ID* row_end, V_ID* src,
                         int32_t* weight) {
  E_ID ne = 2 * ratings;
  V_ID nv = users + items;
  uint64_t* keys = (uint64_t*)malloc(ne * sizeof(uint64_t));
  if (!keys) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < (int64_t)ratings; j++) {
    V_ID u, it;
    lo_bipartite_edge(seed, (uint64_t)j, users, items, &u, &it);
    keys[2 * j] = ((uint64_t)it << 32) | u;     /* user -> item */
    keys[2 * j + 1] = ((uint64_t)u << 32) | it; /* item -> user */
  }
  radix_sort_u64(keys, ne, 64);
  keys_to_csc(keys, ne, nv, row_end, src);
  E_ID e = 0;
  for (V_ID v = 0; v < nv; v++)
    for (; e < row_end[v]; e++) {
      V_ID s = src[e];
      V_ID lo = s < v ? s : v, hi = s < v ? v : s; /* (user,item) regardless of direction */
      weight[e] = lo_edge_weight(seed, lo, hi);
    }
  free(keys);
  return 0;
}

/* Generic: arbitrary edge list -> canonical CSC. */
int lo_edges_to_csc(V_ID nv, E_ID ne, const V_ID* esrc, const V_ID* edst, E_ID* row_end, V_ID* src) {
  uint64_t* keys = (uint64_t*)malloc((ne ? ne : 1) * sizeof(uint64_t));
  if (!keys) return -1;
  for (E_ID i = 0; i < ne; i++) {
    if (esrc[i] >= nv || edst[i] >= nv) { free(keys); return -2; }
    keys[i] = ((uint64_t)edst[i] << 32) | esrc[i];
  }
  radix_sort_u64(keys, ne, 64);
  keys_to_csc(keys, ne, nv, row_end, src);
  free(keys);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * .lux file format — tools/converter.cc:98-124 (writer), pull_model.inl:33-38,294-318 (reader).
 *   u32 nv | u64 ne | u64 row_end[nv] | u32 src[ne] | trailer: i32 weight[ne] (EDGE_WEIGHT apps,
 *   pull_model.inl:309-317) or u32 out_degree[nv] (what converter.cc:124 appends; apps ignore it).
 * ------------------------------------------------------------------------------------------ */
int lo_lux_write(const char* path, V_ID nv, E_ID ne, const E_ID* row_end, const V_ID* src,
                 const int32_t* weight /* may be NULL -> out-degree trailer like converter.cc */) {
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  fwrite(&nv, sizeof(V_ID), 1, f);
  fwrite(&ne, sizeof(E_ID), 1, f);
  fwrite(row_end, sizeof(E_ID), nv, f);
  fwrite(src, sizeof(V_ID), ne, f);
  if (weight) {
    fwrite(weight, sizeof(int32_t), ne, f);
  } else {
    V_ID* deg = (V_ID*)calloc(nv ? nv : 1, sizeof(V_ID));
    for (E_ID e = 0; e < ne; e++) deg[src[e]]++;
    fwrite(deg, sizeof(V_ID), nv, f);
    free(deg);
  }
  fclose(f);
  return 0;
}

int lo_lux_read_header(const char* path, V_ID* nv, E_ID* ne) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  int ok = fread(nv, sizeof(V_ID), 1, f) == 1 && fread(ne, sizeof(E_ID), 1, f) == 1;
  fclose(f);
  return ok ? 0 : -2;
}

int lo_lux_read(const char* path, V_ID nv, E_ID ne, E_ID* row_end, V_ID* src, int32_t* weight /* or NULL */) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  V_ID nv2;
  E_ID ne2;
  if (fread(&nv2, sizeof(V_ID), 1, f) != 1 || fread(&ne2, sizeof(E_ID), 1, f) != 1 || nv2 != nv || ne2 != ne) {
    fclose(f);
    return -2;
  }
  int ok = fread(row_end, sizeof(E_ID), nv, f) == nv && fread(src, sizeof(V_ID), ne, f) == ne;
  if (ok && weight) ok = fread(weight, sizeof(int32_t), ne, f) == ne;
  fclose(f);
  if (!ok) return -3;
  for (V_ID v = 1; v < nv; v++)
    if (row_end[v] < row_end[v - 1]) return -4; /* pull_model.inl:100-101 */
  if (nv && row_end[nv - 1] != ne) return -5;  /* pull_model.inl:102 */
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * A.1 Partitioner — pull_model.inl:108-131 == push_model.inl:378-413.
 * Returns the number of partitions the reference's greedy scan produces (the reference asserts
 * this equals P, pull_model.inl:131).  fq_left/fq_right: frontier-slot byte ranges
 * (push_model.inl:393-397), may be NULL.
 * ------------------------------------------------------------------------------------------ */
int lo_partition(V_ID nv, E_ID ne, const E_ID* row_end, int P, V_ID* row_left, V_ID* row_right, E_ID* col_left,
                 uint64_t* fq_left, uint64_t* fq_right) {
  V_ID left = 0;
  E_ID cnt = 0;
  E_ID cap = (ne + P - 1) / P;
  int count = 0;
  uint64_t fsize = 0;
  for (V_ID v = 0; v < nv; v++) {
    cnt += (v == 0) ? row_end[0] : row_end[v] - row_end[v - 1];
    if (cnt > cap) {
      if (count < P) {
        row_left[count] = left;
        row_right[count] = v;
        if (fq_left) {
          V_ID slots = (v - left) / LO_SPARSE_THRESHOLD + 100;
          fq_left[count] = fsize;
          fsize += 8 + (uint64_t)slots * 4;
          fq_right[count] = fsize - 1;
        }
      }
      count++;
      cnt = 0;
      left = v + 1;
    }
  }
  if (cnt > 0) {
    if (count < P) {
      row_left[count] = left;
      row_right[count] = nv - 1;
      if (fq_left) {
        V_ID slots = (nv - 1 - left) / LO_SPARSE_THRESHOLD + 100;
        fq_left[count] = fsize;
        fsize += 8 + (uint64_t)slots * 4;
        fq_right[count] = fsize - 1;
      }
    }
    count++;
  }
  int n = count < P ? count : P;
  for (int p = 0; p < n; p++) col_left[p] = (row_left[p] == 0) ? 0 : row_end[row_left[p] - 1];
  return count;
}

/* ------------------------------------------------------------------------------------------
 * A.2 PageRank — pull_model.inl:333-343 (out-degree), pagerank_gpu.cu:255-259 (init),
 * pagerank_gpu.cu:86-100,144 (iteration), pagerank/app.h:24 (alpha).
 * Per-vertex sum accumulated in fp64 and rounded once to f32 (SURVEY §8c "PageRank output
 * definition"): the reference's own float atomicAdd order is nondeterministic (pagerank_gpu.cu:90).
 * ------------------------------------------------------------------------------------------ */
void lo_out_degree(V_ID nv, E_ID ne, const V_ID* src, V_ID* deg) {
  memset(deg, 0, (size_t)nv * sizeof(V_ID));
  for (E_ID e = 0; e < ne; e++) deg[src[e]]++;
}

void lo_pagerank_init(V_ID nv, const V_ID* deg, float* x) {
  float rank = 1.0f / nv; /* pagerank_gpu.cu:255 */
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < (int64_t)nv; v++) x[v] = deg[v] == 0 ? rank : rank / deg[v];
}

/* the same iteration over a COMPACT set of destination vertices (lo_rmat_blocks_*): local vertex i has global id
 * vid[i] and in-edges src[row_end_local[i-1] .. row_end_local[i]); x_old / deg are indexed globally, x_new_local by i */
void lo_pagerank_iter_compact(V_ID nv, uint64_t n_local, const E_ID* row_end_local, const V_ID* src, const V_ID* vid,
                              const V_ID* deg, const float* x_old, float* x_new_local) {
  const float init_rank = (1 - LO_ALPHA) / nv; /* pagerank_gpu.cu:144 */
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < (int64_t)n_local; i++) {
    E_ID b = i == 0 ? 0 : row_end_local[i - 1], e = row_end_local[i];
    double s = 0.0;
    for (E_ID k = b; k < e; k++) s += (double)x_old[src[k]];
    float y = fmaf(LO_ALPHA, (float)s, init_rank);
    V_ID dg = deg[vid[i]];
    if (dg != 0) y = y / (float)dg;
    x_new_local[i] = y;
  }
}

/* one iteration over destination vertices [v_lo, v_hi]; x_new indexed globally */
void lo_pagerank_iter_range(V_ID nv, const E_ID* row_end, const V_ID* src, const V_ID* deg, const float* x_old,
                            float* x_new, V_ID v_lo, V_ID v_hi) {
  const float init_rank = (1 - LO_ALPHA) / nv; /* pagerank_gpu.cu:144 */
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t vv = v_lo; vv <= (int64_t)v_hi; vv++) {
    V_ID v = (V_ID)vv;
    E_ID b = v == 0 ? 0 : row_end[v - 1], e = row_end[v];
    double s = 0.0;
    for (E_ID k = b; k < e; k++) s += (double)x_old[src[k]];
    float y = fmaf(LO_ALPHA, (float)s, init_rank); /* :97 (nvcc contracts to FMA) */
    if (deg[v] != 0) y = y / (float)deg[v];       /* :98-99 */
    x_new[v] = y;
  }
}

void lo_pagerank_iter(V_ID nv, const E_ID* row_end, const V_ID* src, const V_ID* deg, const float* x_old,
                      float* x_new) {
  if (nv) lo_pagerank_iter_range(nv, row_end, src, deg, x_old, x_new, 0, nv - 1);
}

/* ------------------------------------------------------------------------------------------
 * A.3 / A.4 CC (max-label) and SSSP (= BFS depth, INF = nv).
 * init: components_gpu.cu:733-739 / sssp_gpu.cu:733-744.
 * pull: components_gpu.cu:112-122 / sssp_gpu.cu:112-122.
 * push: components_gpu.cu:48-82,165-245.   direction rule :414.   frontier rules :408,:462-491.
 * ------------------------------------------------------------------------------------------ */
void lo_label_init(int app, V_ID nv, V_ID start, V_ID* label) {
  if (app == LO_APP_CC)
    for (V_ID v = 0; v < nv; v++) label[v] = v;
  else {
    for (V_ID v = 0; v < nv; v++) label[v] = nv;
    if (start < nv) label[start] = 0;
  }
}

static inline V_ID relax_val(int app, V_ID src_label) { return app == LO_APP_CC ? src_label : src_label + 1; }
static inline int better(int app, V_ID cand, V_ID cur) { return app == LO_APP_CC ? cand > cur : cand < cur; }

/* Jacobi pull sweep over [v_lo, v_hi]; returns #vertices whose label changed. */
uint64_t lo_label_pull_range(int app, V_ID nv, const E_ID* row_end, const V_ID* src, const V_ID* old_l, V_ID* new_l,
                             V_ID v_lo, V_ID v_hi) {
  uint64_t changed = 0;
  (void)nv;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : changed)
  for (int64_t vv = v_lo; vv <= (int64_t)v_hi; vv++) {
    V_ID v = (V_ID)vv;
    E_ID b = v == 0 ? 0 : row_end[v - 1], e = row_end[v];
    V_ID cur = old_l[v];
    for (E_ID k = b; k < e; k++) {
      V_ID c = relax_val(app, old_l[src[k]]);
      if (better(app, c, cur)) cur = c;
    }
    new_l[v] = cur;
    changed += cur != old_l[v];
  }
  return changed;
}

/* CSR-by-source over edges [e_lo, e_hi) whose dsts are [v_lo, v_hi] (one partition's own edges):
 * out_end[u] = END offset (absolute within the partition's edge block) of u's out-edge list,
 * out_dst[] = destination ids.  Restates init_push_kernel / init_push_row_ptrs /
 * init_push_col_idxs (components_gpu.cu:550-607) without defect B3; order inside a source's
 * list is by ascending edge index (the reference's is nondeterministic). */
void lo_build_push_csr(V_ID nv, const E_ID* row_end, const V_ID* src, V_ID v_lo, V_ID v_hi, E_ID* out_end,
                       V_ID* out_dst) {
  E_ID e_lo = v_lo == 0 ? 0 : row_end[v_lo - 1];
  E_ID e_hi = row_end[v_hi];
  E_ID* cursor = (E_ID*)calloc((size_t)nv + 1, sizeof(E_ID));
  for (E_ID e = e_lo; e < e_hi; e++) cursor[src[e] + 1]++;
  for (V_ID u = 0; u < nv; u++) cursor[u + 1] += cursor[u];
  for (V_ID u = 0; u < nv; u++) out_end[u] = cursor[u + 1];
  for (V_ID v = v_lo; v <= v_hi; v++) {
    E_ID b = v == 0 ? 0 : row_end[v - 1];
    for (E_ID e = b; e < row_end[v]; e++) out_dst[cursor[src[e]]++] = v;
  }
  free(cursor);
}

/* Whole-graph run with the reference's iteration structure, P partitions, per-partition frontier
 * slots.  Records per iteration: total active count (Σ numNodes returned by the partitions,
 * components.cc:116-122), direction taken (1 = pull), and each partition's frontier type.
 * Halts at the first iteration that reports 0 active on every partition (the reference's sliding
 * window only delays noticing this, defect B6).  Defect B5 (over-reported numNodes after a
 * sparse->dense promotion) is NOT replicated: counts are exact.
 * Returns number of iterations executed (including the final all-zero one), or <0 on error. */
int lo_label_run(int app, V_ID nv, E_ID ne, const E_ID* row_end, const V_ID* src, int P, V_ID start, V_ID* label_out,
                 int max_iters, uint64_t* active_per_iter, int* pull_per_iter, uint32_t* type_per_iter_part) {
  V_ID *rl = malloc(sizeof(V_ID) * P), *rr = malloc(sizeof(V_ID) * P);
  E_ID* cl = malloc(sizeof(E_ID) * P);
  uint64_t *fl = malloc(sizeof(uint64_t) * P), *fr = malloc(sizeof(uint64_t) * P);
  int np = lo_partition(nv, ne, row_end, P, rl, rr, cl, fl, fr);
  if (np == P - 1 && (np == 0 || rr[np - 1] < nv - 1)) {
    /* the reference would assert here (pull_model.inl:131): the remainder holds only zero-in-degree vertices.
     * Keep them as a last, edge-free partition (same rule as the product) so that no vertex is dropped. */
    rl[np] = np == 0 ? 0 : rr[np - 1] + 1;
    rr[np] = nv - 1;
    cl[np] = ne;
    fl[np] = np == 0 ? 0 : fr[np - 1] + 1;
    fr[np] = fl[np] + 8 + (uint64_t)((rr[np] - rl[np]) / LO_SPARSE_THRESHOLD + 100) * 4 - 1;
    np++;
  }
  if (np != P) { free(rl); free(rr); free(cl); free(fl); free(fr); return -1; }
  V_ID* old_l = malloc(sizeof(V_ID) * (size_t)nv);
  V_ID* new_l = malloc(sizeof(V_ID) * (size_t)nv);
  uint8_t* active = calloc(nv, 1); /* vertex changed in the previous iteration */
  uint32_t* ftype = malloc(sizeof(uint32_t) * P);
  uint64_t* fcount = malloc(sizeof(uint64_t) * P);
  lo_label_init(app, nv, start, new_l);
  /* initial frontier: CC all vertices, dense (components_gpu.cu:733-737); SSSP {start}, sparse */
  for (int p = 0; p < P; p++) {
    if (app == LO_APP_CC) {
      ftype[p] = LO_DENSE_BITMAP;
      fcount[p] = rr[p] - rl[p] + 1;
      for (V_ID v = rl[p]; v <= rr[p]; v++) active[v] = 1;
    } else {
      ftype[p] = LO_SPARSE_QUEUE;
      fcount[p] = (start >= rl[p] && start <= rr[p]) ? 1 : 0;
    }
  }
  if (app == LO_APP_SSSP && start < nv) active[start] = 1;
  /* out-edge index for push steps (global CSR-by-source; union of per-partition ones) */
  E_ID* out_end = malloc(sizeof(E_ID) * (size_t)nv);
  V_ID* out_dst = malloc(sizeof(V_ID) * (size_t)(ne ? ne : 1));
  lo_build_push_csr(nv, row_end, src, 0, nv - 1, out_end, out_dst);

  int it = 0;
  for (; it < max_iters; it++) {
    memcpy(old_l, new_l, sizeof(V_ID) * (size_t)nv); /* components_gpu.cu:391 */
    uint64_t old_size = 0;
    int dense_parts = 0, sparse_parts = 0;
    for (int p = 0; p < P; p++) {
      old_size += fcount[p];
      if (ftype[p] == LO_DENSE_BITMAP) dense_parts++; else sparse_parts++;
    }
    int dense_fq = dense_parts >= sparse_parts; /* :408 */
    int pull = old_size > (uint64_t)(nv / 16);   /* :414 */
    if (pull) {
      dense_fq = 1; /* :416 */
      lo_label_pull_range(app, nv, row_end, src, old_l, new_l, 0, nv - 1);
    } else {
      for (V_ID u = 0; u < nv; u++) {
        if (!active[u]) continue;
        V_ID c = relax_val(app, old_l[u]);
        for (E_ID k = (u == 0 ? 0 : out_end[u - 1]); k < out_end[u]; k++) {
          V_ID d = out_dst[k];
          if (better(app, c, new_l[d])) new_l[d] = c;
        }
      }
    }
    uint64_t total = 0;
    for (int p = 0; p < P; p++) {
      uint64_t c = 0;
      for (V_ID v = rl[p]; v <= rr[p]; v++) { active[v] = old_l[v] != new_l[v]; c += active[v]; }
      uint64_t max_nodes = (fr[p] - fl[p] + 1 - 8) / 4; /* :410 */
      int d = dense_fq;
      if (d) { if (c < max_nodes) d = 0; }   /* demote :469-478 */
      else { if (c >= max_nodes) d = 1; }     /* promote :485-490 */
      ftype[p] = d ? LO_DENSE_BITMAP : LO_SPARSE_QUEUE;
      fcount[p] = c;
      total += c;
      if (type_per_iter_part) type_per_iter_part[(size_t)it * P + p] = ftype[p];
    }
    if (active_per_iter) active_per_iter[it] = total;
    if (pull_per_iter) pull_per_iter[it] = pull;
    if (total == 0) { it++; break; }
  }
  memcpy(label_out, new_l, sizeof(V_ID) * (size_t)nv);
  free(rl); free(rr); free(cl); free(fl); free(fr); free(old_l); free(new_l); free(active); free(ftype);
  free(fcount); free(out_end); free(out_dst);
  return it;
}

/* A.6 invariant checks — components_gpu.cu:768-792, sssp_gpu.cu:773-798. Returns #mistakes. */
uint64_t lo_label_check(int app, V_ID nv, const E_ID* row_end, const V_ID* src, const V_ID* label) {
  uint64_t bad = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : bad)
  for (int64_t vv = 0; vv < (int64_t)nv; vv++) {
    V_ID v = (V_ID)vv;
    for (E_ID k = (v == 0 ? 0 : row_end[v - 1]); k < row_end[v]; k++) {
      V_ID u = src[k];
      if (app == LO_APP_CC) bad += label[v] < label[u];
      else bad += (label[u] != nv) && (label[v] > label[u] + 1);
    }
  }
  return bad;
}

/* ------------------------------------------------------------------------------------------
 * A.5 Collaborative filtering, intended math — col_filter/app.h:26-28, colfilter_gpu.cu:83-100
 * (update), :260-264 (init).  Reference kernel defects are not replicated.
 * Dot product and error in f32 in index order (as the reference's inner loops :84-86); the
 * per-vertex accumulator over edges is fp64 rounded once (same policy as PageRank).
 * ------------------------------------------------------------------------------------------ */
void lo_cf_init(V_ID nv, float* x) {
  float value = sqrtf(1.0f / LO_CF_K);
  for (size_t i = 0; i < (size_t)nv * LO_CF_K; i++) x[i] = value;
}

void lo_cf_iter_range(V_ID nv, const E_ID* row_end, const V_ID* src, const int32_t* w, const float* x_old, float* x_new,
                      V_ID v_lo, V_ID v_hi) {
  (void)nv;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t vv = v_lo; vv <= (int64_t)v_hi; vv++) {
    V_ID v = (V_ID)vv;
    const float* xv = x_old + (size_t)v * LO_CF_K;
    double acc[LO_CF_K];
    for (int i = 0; i < LO_CF_K; i++) acc[i] = 0.0;
    for (E_ID k = (v == 0 ? 0 : row_end[v - 1]); k < row_end[v]; k++) {
      const float* xu = x_old + (size_t)src[k] * LO_CF_K;
      float dot = 0.0f;
      for (int i = 0; i < LO_CF_K; i++) dot = fmaf(xu[i], xv[i], dot);
      float err = (float)w[k] - dot;
      for (int i = 0; i < LO_CF_K; i++) acc[i] += (double)(err * xu[i]);
    }
    for (int i = 0; i < LO_CF_K; i++)
      x_new[(size_t)v * LO_CF_K + i] = xv[i] + LO_CF_GAMMA * ((float)acc[i] - LO_CF_LAMBDA * xv[i]);
  }
}

void lo_cf_iter(V_ID nv, const E_ID* row_end, const V_ID* src, const int32_t* w, const float* x_old, float* x_new) {
  if (nv) lo_cf_iter_range(nv, row_end, src, w, x_old, x_new, 0, nv - 1);
}
