// api.cu — C ABI of libluxb (include/lux_b200.h) and the thin per-rank host runtime behind it.
// Replaces, for the hot path only: Graph::Graph + load/scan/init tasks, pull_app_task_impl / push_app_task_impl,
// the app driver loops and LuxMapper's placement (see the citations in lux_b200.h).
// Product code: nothing here may include, link or call oracle/.
#include <cub/cub.cuh>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <new>

#include "build.cuh"
#include "cf.cuh"
#include "pull.cuh"
#include "panel.cuh"
#include "seg.cuh"
#include "push.cuh"
#include "runtime.cuh"

using namespace luxb;
static const char* const kPhaseName[12] = {"pull_tile", "fixup", "refresh", "rechunk", "barrier", "panel", "combine", "pack+push", "pull/bcast", "", "", ""};
static void pt_mark(luxb_graph* g, int tag) {
  PhaseTimer& pt = g->pt;
  if (!pt.on) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, g->stream);
  pt.ev.push_back(e);
  pt.tag.push_back(tag);
}
static void pt_print(luxb_graph* g) {
  if (g->pt.on && g->pt.cnt) {
    char line[1024];
    int n = snprintf(line, sizeof(line), "[luxb rank %d] phase means over %ld iterations:", g->cfg.rank, g->pt.cnt);
    double sum = 0;
    for (int k = 0; k < 9; ++k) {
      n += snprintf(line + n, sizeof(line) - n, " %s %.3f ms;", kPhaseName[k], g->pt.sum[k] / g->pt.cnt);
      sum += g->pt.sum[k] / g->pt.cnt;
    }
    snprintf(line + n, sizeof(line) - n, " sum %.3f ms\n", sum);
    fputs(line, stderr);  // one write per rank: the ranks' lines do not interleave
  }
}

static void pt_flush(luxb_graph* g) {
  PhaseTimer& pt = g->pt;
  if (!pt.on || pt.ev.empty()) return;
  cudaStreamSynchronize(g->stream);
  for (size_t i = 1; i < pt.ev.size(); ++i) {
    if (pt.tag[i] < 0) continue;
    float ms = 0;
    cudaEventElapsedTime(&ms, pt.ev[i - 1], pt.ev[i]);
    pt.sum[pt.tag[i]] += ms;
    if (pt.tag[i] == 0) pt.cnt++;
  }
  for (cudaEvent_t e : pt.ev) cudaEventDestroy(e);
  pt.ev.clear();
  pt.tag.clear();
}


// ------------------------------------------------------------------------------------------------------------
namespace luxb {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace luxb

#define LUXB_ARG(cond, ...)             \
  do {                                  \
    if (!(cond)) {                      \
      set_error(__VA_ARGS__);           \
      return LUXB_ERR_ARG;              \
    }                                   \
  } while (0)

#define LUXB_NCCL(expr)                                                                              \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != ncclSuccess) {                                                                         \
      set_error("NCCL error %s at %s:%d: %s", #expr, __FILE__, __LINE__, nccl().GetErrorString(_r)); \
      return LUXB_ERR_COMM;                                                                          \
    }                                                                                                \
  } while (0)

// merge-path shapes <items per lane, consumer warps per CTA, ring stages>; chosen at open time (LUXB_PULL_SHAPE).
// Shared memory is kept small on purpose: what the ring does not take stays L1, and the gather rate follows L1 size.
#define LUXB_PULL_SHAPES(X) X(0, 7, 8, 2) X(1, 9, 8, 2) X(2, 7, 16, 2)
#define LUXB_DECL_SHAPE(id, ipt, warps, stages) using PullShape##id = PullShape<ipt, warps, stages>;
LUXB_PULL_SHAPES(LUXB_DECL_SHAPE)
#define LUXB_TILE_OF(id, ipt, warps, stages) PullShape##id::kTile,
static const int kPullTileOf[] = {LUXB_PULL_SHAPES(LUXB_TILE_OF)};
static const int kNumPullShapes = sizeof(kPullTileOf) / sizeof(int);
static const int kDefaultPullShape = 0;
static const int kDefaultPullCtas = 3;

static inline int grid_for(uint64_t n, int threads, int cap) {
  uint64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (uint64_t)cap) b = cap;
  return (int)b;
}

template <class T>
static int dmalloc(T** p, uint64_t count) {
  void* q = nullptr;
  LUXB_CUDA(cudaMalloc(&q, std::max<uint64_t>(count, 1) * sizeof(T) + 256));
  *p = reinterpret_cast<T*>(q);
  return 0;
}

// edge arrays: HBM, or mapped pinned host memory when cfg.zero_copy_edges (TMA bulk copies and plain loads read it
// over PCIe through the same unified addresses)
template <class T>
static int edge_alloc(luxb_graph* g, T** p, uint64_t count) {
  if (!g->cfg.zero_copy_edges) return dmalloc(p, count);
  void* q = nullptr;
  LUXB_CUDA(cudaHostAlloc(&q, std::max<uint64_t>(count, 1) * sizeof(T) + 256, cudaHostAllocMapped | cudaHostAllocPortable));
  g->host_allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

// ---- the reference partitioner on the host (pull_model.inl:108-131); same cut rule as partition_kernel -------
static int host_partition(uint32_t nv, uint64_t ne, const uint64_t* row_end, int P, uint32_t* rl, uint32_t* np,
                          uint64_t* cl) {
  uint64_t cap = (ne + P - 1) / P;
  uint32_t left = 0;
  int count = 0, found_adjust = 0;
  while (left < nv && count < P) {
    uint64_t base = left == 0 ? 0 : row_end[left - 1];
    const uint64_t* first = std::upper_bound(row_end + left, row_end + nv, base + cap);  // first row_end[v] > base+cap
    uint32_t v = (uint32_t)(first - row_end);
    if (v < nv) {
      rl[count] = left; np[count] = v - left + 1; cl[count] = base; ++count;
      left = v + 1;
    } else {
      // the reference emits the remainder only if it holds edges (pull_model.inl:128-130) and would then assert
      // on the partition count; we always keep trailing zero-in-degree vertices so that none is dropped
      if (row_end[nv - 1] - base == 0) --found_adjust;
      rl[count] = left; np[count] = nv - left; cl[count] = base; ++count;
      left = nv;
    }
  }
  int found = count + found_adjust;
  for (int p = count; p < P; ++p) { rl[p] = nv; np[p] = 0; cl[p] = ne; }
  return found;
}

// ---- cost-balanced work split (cfg.balanced_split, pull apps) ------------------------------------------------------
// The reference's split balances EDGES; the sweep's cost per edge is not uniform (measured at RMAT-27: an edge into a
// hub destination mostly travels through the shared-memory panel, ~1.5-2 ns; any other edge goes through L1, ~3.5 ns;
// every vertex costs ~8 ns of bookkeeping — at 2 GPUs the reference split leaves the ranks 25 % apart, at 8 GPUs rank 7
// owns 40 % of the vertices).  Contiguous destination ranges are kept; only the cut points move.  Weights in integer
// units: vertex 8, edge into a hub (in-degree >= kBalanceHubIndeg) 4, other edge 7 (at 8 GPUs with vertex = 16 the
// vertex-rich last rank finished its sweep a third earlier than rank 0: profiles/r02a_bench_n8_phases.txt).
static constexpr uint32_t kBalanceHubIndeg = 64;
static inline uint64_t vertex_cost(uint64_t indeg) { return 8 + (indeg >= kBalanceHubIndeg ? 4 : 7) * indeg; }

static void host_balanced_partition(uint32_t nv, uint64_t ne, const uint64_t* row_end, int P, uint32_t* rl, uint32_t* np, uint64_t* cl) {
  uint64_t total = 0;
  for (uint32_t v = 0; v < nv; ++v) total += vertex_cost(row_end[v] - (v ? row_end[v - 1] : 0));
  uint64_t run = 0;
  uint32_t left = 0;
  int p = 0;
  for (uint32_t v = 0; v < nv && p < P - 1; ++v) {
    run += vertex_cost(row_end[v] - (v ? row_end[v - 1] : 0));
    if (run * P >= total * (uint64_t)(p + 1)) {  // close partition p at v (inclusive)
      rl[p] = left; np[p] = v - left + 1; cl[p] = left ? row_end[left - 1] : 0;
      left = v + 1;
      ++p;
    }
  }
  rl[p] = left; np[p] = nv - left; cl[p] = left ? row_end[left - 1] : 0;
  for (++p; p < P; ++p) { rl[p] = nv; np[p] = 0; cl[p] = ne; }
}

static bool use_balanced_split(const luxb_config* cfg) {
  return cfg->balanced_split && cfg->nranks > 1 && (cfg->app == LUXB_PAGERANK || cfg->app == LUXB_COLFILTER);
}

static int check_config(const luxb_config* cfg) {
  LUXB_ARG(cfg != nullptr, "config is NULL");
  LUXB_ARG(cfg->app >= LUXB_PAGERANK && cfg->app <= LUXB_COLFILTER, "unknown app %d", (int)cfg->app);
  LUXB_ARG(cfg->nranks >= 1 && cfg->nranks <= LUXB_MAX_PARTS, "nranks %d out of range [1,%d]", cfg->nranks, LUXB_MAX_PARTS);
  LUXB_ARG(cfg->rank >= 0 && cfg->rank < cfg->nranks, "rank %d out of range", cfg->rank);
  return 0;
}

static int graph_begin(const luxb_config* cfg, luxb_graph** out) {
  LUXB_TRY(check_config(cfg));
  LUXB_ARG(out != nullptr, "out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s): libluxb has no CPU fallback", cudaGetErrorString(e));
    return LUXB_ERR_CUDA;
  }
  LUXB_ARG(cfg->device >= 0 && cfg->device < ndev, "device %d out of range (have %d)", cfg->device, ndev);
  LUXB_CUDA(cudaSetDevice(cfg->device));
  luxb_graph* g = new (std::nothrow) luxb_graph();
  if (!g) { set_error("out of host memory"); return LUXB_ERR_NOMEM; }
  g->cfg = *cfg;
  g->P = cfg->nranks;
  *out = g;
  LUXB_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
  LUXB_CUDA(cudaEventCreate(&g->ev_begin));
  LUXB_CUDA(cudaEventCreate(&g->ev_end));
  LUXB_CUDA(cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
  g->pull_ctas = kDefaultPullCtas;
  if (const char* env = getenv("LUXB_PULL_CTAS")) g->pull_ctas = std::max(1, atoi(env));
  if (const char* env = getenv("LUXB_PHASE_TIMING")) { g->pt.on = atoi(env) != 0; g->pt.per_call = atoi(env) == 2; }
  if (const char* env = getenv("LUXB_L2_HINTS")) g->l2_hints = atoi(env);
  if (const char* env = getenv("LUXB_OVERLAP")) g->overlap_exchange = atoi(env) != 0;
  if (const char* env = getenv("LUXB_BARRIER")) { g->flag_barrier = strcmp(env, "nccl") != 0; g->flag_barrier_all = strcmp(env, "flag") == 0; }
  if (const char* env = getenv("LUXB_PUSH")) g->direct_push = strcmp(env, "direct") == 0;
  if (const char* env = getenv("LUXB_BARRIER_TIMEOUT_S")) g->barrier_timeout_ns = (uint64_t)std::max(1, atoi(env)) * 1000000000ull;
  if (const char* env = getenv("LUXB_FUSED_FIXUP")) g->fused_fixup = atoi(env) != 0;
  if (const char* env = getenv("LUXB_PANEL_RESERVE_SMS")) g->panel_reserve_sms = std::max(0, atoi(env));
  return 0;
}

static void set_partition_derived(luxb_graph* g) {
  int r = g->cfg.rank;
  g->row_left = g->rl[r];
  g->n_part = g->np[r];
  g->col_left = g->cl[r];
  uint64_t next = (r + 1 < g->P) ? g->cl[r + 1] : g->ne;
  if (g->np[r] == 0) next = g->col_left;
  g->e_part = next - g->col_left;
  uint64_t off = 0;
  for (int p = 0; p < g->P; ++p) {
    uint32_t span = g->np[p] ? g->np[p] - 1 : 0;          // R - L
    g->cap[p] = span / 16 + 100;                          // push_model.inl:393
    g->slot_off[p] = off;
    g->slot_bytes[p] = ((8 + (uint64_t)g->cap[p] * 8) + 15) & ~15ull;  // header + ids + labels
    off += g->slot_bytes[p];
  }
  g->fq_total = off;
}

// after d_row_end / d_src are in place: merge-path tile table + per-tile partial buffers
static int finish_layout(luxb_graph* g) {
  uint64_t total = (uint64_t)g->n_part + g->e_part;
  g->pull_shape = kDefaultPullShape;
  if (const char* env = getenv("LUXB_PULL_SHAPE")) {
    int v = atoi(env);
    if (v >= 0 && v < kNumPullShapes) g->pull_shape = v;
  }
  const uint32_t tile = (uint32_t)kPullTileOf[g->pull_shape];
  uint64_t nt = (total + tile - 1) / tile;
  LUXB_ARG(nt < 0xFFFFFFFFull, "partition too large for the tile table");
  g->n_tiles = (uint32_t)nt;
  LUXB_TRY(dmalloc(&g->d_tile_v, (uint64_t)g->n_tiles + 2));
  tile_table_kernel<<<grid_for((uint64_t)g->n_tiles + 1, 256, 1 << 20), 256, 0, g->stream>>>(
      g->d_row_end, g->n_part, g->e_part, tile, g->n_tiles, g->d_tile_v);
  LUXB_CUDA(cudaGetLastError());
  LUXB_TRY(dmalloc(&g->d_row_end32, (uint64_t)g->n_part + 8));
  narrow_u64_to_u32_kernel<<<grid_for((uint64_t)g->n_part + 8, 256, 4096), 256, 0, g->stream>>>(g->d_row_end, g->n_part + 4,
                                                                                          g->d_row_end32, (uint64_t)g->n_part + 8);
  LUXB_CUDA(cudaGetLastError());
  LUXB_TRY(dmalloc((uint32_t**)&g->d_head, (uint64_t)g->n_tiles + 1));
  LUXB_TRY(dmalloc((uint32_t**)&g->d_tail, (uint64_t)g->n_tiles + 1));
  g->n_fix_blocks = (g->n_tiles + kFixBlock - 1) / kFixBlock;
  LUXB_TRY(dmalloc((uint64_t**)&g->d_carry, (uint64_t)g->n_tiles + 1));
  LUXB_TRY(dmalloc(&g->d_carry_flag, (uint64_t)g->n_tiles + 1));
  LUXB_TRY(dmalloc((uint64_t**)&g->d_block_agg, (uint64_t)g->n_fix_blocks + 1));
  LUXB_TRY(dmalloc(&g->d_block_flag, (uint64_t)g->n_fix_blocks + 1));
  LUXB_TRY(dmalloc(&g->d_counters, 8));  // [0] edges scanned, [1] check mistakes, [2] pull tile counter, [3] big segments, [4] panel tile counter
  LUXB_CUDA(cudaMemsetAsync(g->d_counters, 0, 8 * sizeof(unsigned long long), g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  return 0;
}

static int validate_row_end(uint32_t nv, uint64_t ne, const uint64_t* row_end) {
  LUXB_ARG(nv >= 1, "graph has no vertices");
  // device-wide scans / sorts index vertices with 32-bit signed counts and INF = nv must stay a valid label
  LUXB_ARG(nv < 0x7FFFFFFEu, "nv = %u: at most 2^31 - 2 vertices are supported", nv);
  for (uint32_t v = 1; v < nv; ++v)
    LUXB_ARG(row_end[v] >= row_end[v - 1], "row_end not non-decreasing at vertex %u (pull_model.inl:100-101)", v);
  LUXB_ARG(row_end[nv - 1] == ne, "row_end[nv-1] (%llu) != ne (%llu) (pull_model.inl:102)",
           (unsigned long long)row_end[nv - 1], (unsigned long long)ne);
  return 0;
}

// upload this rank's slice given host pointers to ITS portion of row_end (absolute) / src / weight
static int upload_slice(luxb_graph* g, const uint64_t* row_end_slice_abs, const uint32_t* src_slice,
                        const int32_t* weight_slice) {
  LUXB_TRY(dmalloc(&g->d_row_end, (uint64_t)g->n_part + 4));
  LUXB_TRY(edge_alloc(g, &g->d_src, g->e_part + 8));
  uint64_t* d_tmp = nullptr;
  LUXB_TRY(dmalloc(&d_tmp, (uint64_t)g->n_part + 1));
  if (g->n_part)
    LUXB_CUDA(cudaMemcpyAsync(d_tmp, row_end_slice_abs, (size_t)g->n_part * 8, cudaMemcpyHostToDevice, g->stream));
  rowend_rel_kernel<<<grid_for((uint64_t)g->n_part + 4, 256, 4096), 256, 0, g->stream>>>(d_tmp, 0, g->n_part, g->col_left,
                                                                                        g->d_row_end);
  LUXB_CUDA(cudaGetLastError());
  LUXB_CUDA(cudaMemsetAsync(g->d_src, 0, (g->e_part + 8) * 4, g->stream));
  if (g->e_part) LUXB_CUDA(cudaMemcpyAsync(g->d_src, src_slice, g->e_part * 4, cudaMemcpyDefault, g->stream));
  if (g->weighted) {
    LUXB_TRY(edge_alloc(g, &g->d_weight, g->e_part + 8));
    LUXB_CUDA(cudaMemsetAsync(g->d_weight, 0, (g->e_part + 8) * 4, g->stream));
    if (g->e_part) LUXB_CUDA(cudaMemcpyAsync(g->d_weight, weight_slice, g->e_part * 4, cudaMemcpyDefault, g->stream));
  }
  // every source id of this rank's slice must be a vertex (checked where the data already is: on the device)
  unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(d_tmp);
  LUXB_CUDA(cudaMemsetAsync(d_bad, 0, 8, g->stream));
  if (g->e_part) src_out_of_range_kernel<<<g->num_sms * 8, 256, 0, g->stream>>>(g->d_src, g->e_part, g->nv, d_bad);
  unsigned long long bad = 0;
  LUXB_CUDA(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  LUXB_CUDA(cudaFree(d_tmp));
  LUXB_ARG(bad == 0, "%llu source ids of this rank's slice are >= nv (%u)", bad, g->nv);
  return finish_layout(g);
}

// ------------------------------------------------------------------------------------------------------------
extern "C" {

const char* luxb_last_error(void) { return g_err; }
const char* luxb_version(void) { return "lux_b200 0.2 (sm_100a)"; }
int luxb_abi_version(void) { return 3; }

int luxb_partition_csc(luxb_vid nv, luxb_eid ne, const luxb_eid* row_end, int P, luxb_vid* row_left, luxb_vid* row_right,
                       luxb_eid* col_left) {
  LUXB_ARG(row_end && row_left && row_right && col_left, "NULL argument");
  LUXB_ARG(P >= 1 && P <= LUXB_MAX_PARTS, "P out of range");
  LUXB_TRY(validate_row_end(nv, ne, row_end));
  uint32_t np[LUXB_MAX_PARTS];
  int found = host_partition(nv, ne, row_end, P, row_left, np, col_left);
  for (int p = 0; p < P; ++p) row_right[p] = row_left[p] + np[p] - 1;  // empty: row_left - 1
  return found;
}

int luxb_open_csc(const luxb_csc* csc, const luxb_config* cfg, luxb_graph** out) {
  LUXB_ARG(csc && csc->row_end && (csc->src || csc->ne == 0), "csc arrays are NULL");
  LUXB_TRY(check_config(cfg));
  LUXB_ARG(cfg->app != LUXB_COLFILTER || csc->weight, "col_filter needs edge weights (EDGE_WEIGHT, col_filter/app.h:22)");
  LUXB_TRY(validate_row_end(csc->nv, csc->ne, csc->row_end));
  luxb_graph* g = nullptr;
  int rc = graph_begin(cfg, &g);
  if (rc) { if (g) luxb_close(g); return rc; }
  g->nv = csc->nv;
  g->ne = csc->ne;
  g->weighted = cfg->app == LUXB_COLFILTER;
  g->parts_found = host_partition(g->nv, g->ne, csc->row_end, g->P, g->ref_rl, g->ref_np, g->ref_cl);
  if (use_balanced_split(cfg)) host_balanced_partition(g->nv, g->ne, csc->row_end, g->P, g->rl, g->np, g->cl);
  else for (int p = 0; p < g->P; ++p) { g->rl[p] = g->ref_rl[p]; g->np[p] = g->ref_np[p]; g->cl[p] = g->ref_cl[p]; }
  set_partition_derived(g);
  rc = upload_slice(g, csc->row_end + (g->n_part ? g->row_left : 0), csc->src + g->col_left,
                    g->weighted ? csc->weight + g->col_left : nullptr);
  if (rc) { luxb_close(g); return rc; }
  *out = g;
  return 0;
}

int luxb_open_file(const char* path, const luxb_config* cfg, luxb_graph** out) {
  LUXB_ARG(path != nullptr, "path is NULL");
  LUXB_TRY(check_config(cfg));
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("cannot open %s", path); return LUXB_ERR_IO; }
  uint32_t nv = 0;
  uint64_t ne = 0;
  if (fread(&nv, 4, 1, f) != 1 || fread(&ne, 8, 1, f) != 1 || nv == 0) {  // FILE_HEADER_SIZE, core/graph.h:32
    fclose(f);
    set_error("%s: bad .lux header", path);
    return LUXB_ERR_IO;
  }
  std::vector<uint64_t> row_end(nv);
  if (fread(row_end.data(), 8, nv, f) != nv) { fclose(f); set_error("%s: truncated row_end", path); return LUXB_ERR_IO; }
  int rc = validate_row_end(nv, ne, row_end.data());
  if (rc) { fclose(f); return rc; }
  luxb_graph* g = nullptr;
  rc = graph_begin(cfg, &g);
  if (rc) { fclose(f); if (g) luxb_close(g); return rc; }
  g->nv = nv;
  g->ne = ne;
  g->weighted = cfg->app == LUXB_COLFILTER;
  g->parts_found = host_partition(nv, ne, row_end.data(), g->P, g->ref_rl, g->ref_np, g->ref_cl);
  if (use_balanced_split(cfg)) host_balanced_partition(nv, ne, row_end.data(), g->P, g->rl, g->np, g->cl);
  else for (int p = 0; p < g->P; ++p) { g->rl[p] = g->ref_rl[p]; g->np[p] = g->ref_np[p]; g->cl[p] = g->ref_cl[p]; }
  set_partition_derived(g);
  // this rank's slice only — same seeks as pull_load_task_impl (pull_model.inl:294-318)
  std::vector<uint32_t> src(g->e_part ? g->e_part : 1);
  std::vector<int32_t> w(g->weighted && g->e_part ? g->e_part : 1);
  bool ok = fseeko(f, (off_t)(12 + 8 * (uint64_t)nv + 4 * g->col_left), SEEK_SET) == 0 &&
            fread(src.data(), 4, g->e_part, f) == g->e_part;
  if (ok && g->weighted)
    ok = fseeko(f, (off_t)(12 + 8 * (uint64_t)nv + 4 * ne + 4 * g->col_left), SEEK_SET) == 0 &&
         fread(w.data(), 4, g->e_part, f) == g->e_part;
  fclose(f);
  if (!ok) { luxb_close(g); set_error("%s: truncated edge data", path); return LUXB_ERR_IO; }
  rc = upload_slice(g, row_end.data() + (g->n_part ? g->row_left : 0), src.data(), g->weighted ? w.data() : nullptr);
  if (rc) { luxb_close(g); return rc; }
  *out = g;
  return 0;
}

// ---- .lux writer and edge-list converter (tools/converter.cc) — host only, no device needed -----------------------
int luxb_write_lux(const char* path, const luxb_csc* csc) {
  LUXB_ARG(path && csc && csc->row_end && (csc->src || csc->ne == 0), "NULL argument");
  LUXB_TRY(validate_row_end(csc->nv, csc->ne, csc->row_end));
  std::vector<uint32_t> deg;
  if (!csc->weight) {  // the trailer the reference converter writes: out-degrees (converter.cc:124)
    deg.assign(csc->nv, 0);
    for (uint64_t e = 0; e < csc->ne; ++e) {
      LUXB_ARG(csc->src[e] < csc->nv, "src[%llu] out of range", (unsigned long long)e);
      deg[csc->src[e]]++;
    }
  }
  FILE* f = fopen(path, "wb");
  if (!f) { set_error("cannot create %s", path); return LUXB_ERR_IO; }
  bool ok = fwrite(&csc->nv, 4, 1, f) == 1 && fwrite(&csc->ne, 8, 1, f) == 1 &&       // converter.cc:108-109
            fwrite(csc->row_end, 8, csc->nv, f) == csc->nv &&                          // :110
            (csc->ne == 0 || fwrite(csc->src, 4, csc->ne, f) == csc->ne);              // :112-123
  if (ok && csc->weight) ok = csc->ne == 0 || fwrite(csc->weight, 4, csc->ne, f) == csc->ne;  // EDGE_WEIGHT apps read i32 weights here
  if (ok && !csc->weight) ok = fwrite(deg.data(), 4, csc->nv, f) == csc->nv;                  // :124
  ok = fclose(f) == 0 && ok;
  if (!ok) { set_error("short write to %s", path); return LUXB_ERR_IO; }
  return 0;
}

int luxb_convert_edgelist(const char* edge_list_path, const char* lux_path, luxb_vid nv, luxb_eid ne) {
  LUXB_ARG(edge_list_path && lux_path, "NULL argument");
  LUXB_ARG(nv >= 1, "-nv must be positive");
  FILE* fin = fopen(edge_list_path, "r");
  if (!fin) { set_error("cannot open %s", edge_list_path); return LUXB_ERR_IO; }
  std::vector<uint64_t> keys;  // dst << 32 | src: ascending = canonical (dst, src) order (the reference's std::sort by dst
  keys.reserve(ne);            // leaves the order inside a destination unspecified; ours is deterministic)
  for (uint64_t e = 0; e < ne; ++e) {
    long long a = -1, b = -1;
    if (fscanf(fin, "%lli %lli", &a, &b) != 2) {  // "%i %i" in converter.cc:90: C integer syntax, whitespace separated
      fclose(fin);
      set_error("%s: edge %llu of %llu cannot be read", edge_list_path, (unsigned long long)e, (unsigned long long)ne);
      return LUXB_ERR_IO;
    }
    if (a < 0 || b < 0 || (unsigned long long)a >= nv || (unsigned long long)b >= nv) {  // converter.cc:91-92 asserts
      fclose(fin);
      set_error("%s: edge %llu (%lld -> %lld) has an endpoint outside [0, %u)", edge_list_path, (unsigned long long)e, a, b, nv);
      return LUXB_ERR_ARG;
    }
    keys.push_back(((uint64_t)b << 32) | (uint64_t)a);
  }
  fclose(fin);
  std::sort(keys.begin(), keys.end());
  std::vector<uint64_t> row_end(nv);
  std::vector<uint32_t> src(ne ? ne : 1);
  uint64_t cnt = 0;
  for (uint32_t v = 0; v < nv; ++v) {
    while (cnt < ne && (uint32_t)(keys[cnt] >> 32) == v) { src[cnt] = (uint32_t)keys[cnt]; ++cnt; }
    row_end[v] = cnt;  // END offset of v's in-edge block (converter.cc:100-106)
  }
  luxb_csc csc{nv, ne, row_end.data(), src.data(), nullptr};
  return luxb_write_lux(lux_path, &csc);
}

static int open_generated(const GenSpec& spec, const luxb_config* cfg, luxb_graph** out) {
  luxb_graph* g = nullptr;
  int rc = graph_begin(cfg, &g);
  if (rc) { if (g) luxb_close(g); return rc; }
  auto fail = [&](int code) { luxb_close(g); return code; };
  g->nv = spec.nv;
  g->ne = spec.ne;
  g->weighted = cfg->app == LUXB_COLFILTER;
  const int gen_grid = g->num_sms * 16;
  // 1. in-degree histogram over the whole edge stream -> global row_end (u64) by an inclusive scan
  uint32_t* d_indeg = nullptr;
  uint64_t* d_row_end_g = nullptr;
  if ((rc = dmalloc(&d_indeg, spec.nv))) return fail(rc);
  if ((rc = dmalloc(&d_row_end_g, spec.nv))) return fail(rc);
#define GEN_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { set_error("CUDA error %s: %s", #x, cudaGetErrorString(_e)); return fail(LUXB_ERR_CUDA); } } while (0)
  GEN_CUDA(cudaMemsetAsync(d_indeg, 0, (size_t)spec.nv * 4, g->stream));
  gen_count_indeg_kernel<<<gen_grid, 256, 0, g->stream>>>(spec, d_indeg);
  widen_u32_to_u64_kernel<<<gen_grid, 256, 0, g->stream>>>(d_indeg, d_row_end_g, spec.nv);
  size_t tmp_bytes = 0;
  GEN_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, d_row_end_g, d_row_end_g, (int)spec.nv, g->stream));
  void* d_tmp = nullptr;
  GEN_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 256));
  GEN_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, d_row_end_g, d_row_end_g, (int)spec.nv, g->stream));
  // 2. partition table (reference greedy split), on the device
  uint32_t* d_pt = nullptr;  // rl[P], np[P]
  uint64_t* d_cl = nullptr;
  int* d_cnt = nullptr;
  if ((rc = dmalloc(&d_pt, 2 * LUXB_MAX_PARTS))) return fail(rc);
  if ((rc = dmalloc(&d_cl, LUXB_MAX_PARTS))) return fail(rc);
  if ((rc = dmalloc(&d_cnt, 1))) return fail(rc);
  partition_kernel<<<1, 1, 0, g->stream>>>(d_row_end_g, spec.nv, spec.ne, g->P, d_pt, d_pt + LUXB_MAX_PARTS, d_cl, d_cnt);
  GEN_CUDA(cudaMemcpyAsync(g->ref_rl, d_pt, g->P * 4, cudaMemcpyDeviceToHost, g->stream));
  GEN_CUDA(cudaMemcpyAsync(g->ref_np, d_pt + LUXB_MAX_PARTS, g->P * 4, cudaMemcpyDeviceToHost, g->stream));
  GEN_CUDA(cudaMemcpyAsync(g->ref_cl, d_cl, g->P * 8, cudaMemcpyDeviceToHost, g->stream));
  GEN_CUDA(cudaMemcpyAsync(&g->parts_found, d_cnt, 4, cudaMemcpyDeviceToHost, g->stream));
  GEN_CUDA(cudaStreamSynchronize(g->stream));
  if (use_balanced_split(cfg)) {
    // cost prefix over all vertices (in place of the in-degree scratch), then P - 1 binary searches for the cut points
    uint64_t* d_cost = nullptr;
    if ((rc = dmalloc(&d_cost, (uint64_t)spec.nv + 1))) return fail(rc);
    vertex_cost_kernel<<<gen_grid, 256, 0, g->stream>>>(d_row_end_g, spec.nv, kBalanceHubIndeg, d_cost);
    size_t tb2 = 0;
    GEN_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, d_cost, d_cost, (int)spec.nv, g->stream));
    void* d_tmp2 = nullptr;
    GEN_CUDA(cudaMalloc(&d_tmp2, tb2 + 256));
    GEN_CUDA(cub::DeviceScan::InclusiveSum(d_tmp2, tb2, d_cost, d_cost, (int)spec.nv, g->stream));
    balanced_cut_kernel<<<1, 1, 0, g->stream>>>(d_cost, d_row_end_g, spec.nv, spec.ne, g->P, d_pt, d_pt + LUXB_MAX_PARTS, d_cl);
    GEN_CUDA(cudaMemcpyAsync(g->rl, d_pt, g->P * 4, cudaMemcpyDeviceToHost, g->stream));
    GEN_CUDA(cudaMemcpyAsync(g->np, d_pt + LUXB_MAX_PARTS, g->P * 4, cudaMemcpyDeviceToHost, g->stream));
    GEN_CUDA(cudaMemcpyAsync(g->cl, d_cl, g->P * 8, cudaMemcpyDeviceToHost, g->stream));
    GEN_CUDA(cudaStreamSynchronize(g->stream));
    GEN_CUDA(cudaFree(d_tmp2));
    GEN_CUDA(cudaFree(d_cost));
  } else {
    for (int p = 0; p < g->P; ++p) { g->rl[p] = g->ref_rl[p]; g->np[p] = g->ref_np[p]; g->cl[p] = g->ref_cl[p]; }
  }
  set_partition_derived(g);
  // 3. local row_end (relative + sentinels)
  if ((rc = dmalloc(&g->d_row_end, (uint64_t)g->n_part + 4))) return fail(rc);
  rowend_rel_kernel<<<grid_for((uint64_t)g->n_part + 4, 256, 4096), 256, 0, g->stream>>>(d_row_end_g, g->row_left, g->n_part,
                                                                                        g->col_left, g->d_row_end);
  GEN_CUDA(cudaGetLastError());
  GEN_CUDA(cudaStreamSynchronize(g->stream));
  GEN_CUDA(cudaFree(d_indeg));
  GEN_CUDA(cudaFree(d_row_end_g));
  GEN_CUDA(cudaFree(d_tmp));
  GEN_CUDA(cudaFree(d_pt));
  GEN_CUDA(cudaFree(d_cl));
  GEN_CUDA(cudaFree(d_cnt));
  // 4. this partition's edges: regenerate the stream, keep keys (dst_local << 32 | src), radix sort -> canonical CSC
  uint64_t *d_keys = nullptr, *d_keys_alt = nullptr;
  unsigned long long* d_cursor = nullptr;
  if ((rc = dmalloc(&d_keys, g->e_part))) return fail(rc);
  if ((rc = dmalloc(&d_keys_alt, g->e_part))) return fail(rc);
  if ((rc = dmalloc(&d_cursor, 1))) return fail(rc);
  GEN_CUDA(cudaMemsetAsync(d_cursor, 0, 8, g->stream));
  gen_emit_keys_kernel<<<gen_grid, 256, 0, g->stream>>>(spec, g->row_left, g->n_part, d_cursor, d_keys, g->e_part);
  GEN_CUDA(cudaGetLastError());
  unsigned long long emitted = 0;
  GEN_CUDA(cudaMemcpyAsync(&emitted, d_cursor, 8, cudaMemcpyDeviceToHost, g->stream));
  GEN_CUDA(cudaStreamSynchronize(g->stream));
  if (emitted != g->e_part) {
    set_error("generator emitted %llu edges for this partition, expected %llu", emitted, (unsigned long long)g->e_part);
    return fail(LUXB_ERR_STATE);
  }
  int vbits = 1;
  while ((1ull << vbits) < (uint64_t)spec.nv) ++vbits;
  int pbits = 1;
  while ((1ull << pbits) < (uint64_t)g->n_part + 1) ++pbits;
  cub::DoubleBuffer<uint64_t> keys(d_keys, d_keys_alt);
  tmp_bytes = 0;
  GEN_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, (long long)g->e_part, 0, 32 + pbits, g->stream));
  GEN_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 256));
  GEN_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, keys, (long long)g->e_part, 0, 32 + pbits, g->stream));
  if ((rc = edge_alloc(g, &g->d_src, g->e_part + 8))) return fail(rc);
  GEN_CUDA(cudaMemsetAsync(g->d_src, 0, (g->e_part + 8) * 4, g->stream));
  if (g->weighted) {
    if ((rc = edge_alloc(g, &g->d_weight, g->e_part + 8))) return fail(rc);
    GEN_CUDA(cudaMemsetAsync(g->d_weight, 0, (g->e_part + 8) * 4, g->stream));
  }
  keys_to_src_kernel<<<gen_grid, 256, 0, g->stream>>>(keys.Current(), g->e_part, g->d_src, g->d_weight, spec.seed, g->row_left);
  GEN_CUDA(cudaGetLastError());
  GEN_CUDA(cudaStreamSynchronize(g->stream));
  GEN_CUDA(cudaFree(d_keys));
  GEN_CUDA(cudaFree(d_keys_alt));
  GEN_CUDA(cudaFree(d_cursor));
  GEN_CUDA(cudaFree(d_tmp));
#undef GEN_CUDA
  (void)vbits;
  rc = finish_layout(g);
  if (rc) return fail(rc);
  *out = g;
  return 0;
}

int luxb_open_rmat(int scale, luxb_vid nv, luxb_eid ne, uint64_t seed, const luxb_config* cfg, luxb_graph** out) {
  LUXB_TRY(check_config(cfg));
  LUXB_ARG(scale >= 1 && scale <= 31, "scale out of range");
  LUXB_ARG(nv >= 1 && (uint64_t)nv <= (1ull << scale) && nv < 0x7FFFFFFFu, "nv must be in [1, min(2^scale, 2^31 - 2)]");
  LUXB_ARG(cfg->app != LUXB_COLFILTER, "use luxb_open_bipartite for col_filter");
  GenSpec s{};
  s.kind = 0; s.scale = scale; s.nv = nv; s.ne = ne; s.seed = seed;
  return open_generated(s, cfg, out);
}

int luxb_open_bipartite(luxb_vid users, luxb_vid items, luxb_eid ratings, uint64_t seed, const luxb_config* cfg,
                        luxb_graph** out) {
  LUXB_TRY(check_config(cfg));
  LUXB_ARG(users >= 1 && items >= 1, "users and items must be positive");
  LUXB_ARG((uint64_t)users + items < 0x7FFFFFFEull, "users + items must stay below 2^31 - 2");
  GenSpec s{};
  s.kind = 1; s.nv = users + items; s.ne = 2 * ratings; s.seed = seed; s.users = users; s.items = items;
  return open_generated(s, cfg, out);
}

int luxb_graph_info(const luxb_graph* g, luxb_vid* nv, luxb_eid* ne, int* nranks) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  if (nv) *nv = g->nv;
  if (ne) *ne = g->ne;
  if (nranks) *nranks = g->P;
  return 0;
}

int luxb_partition_bounds(const luxb_graph* g, luxb_vid* row_left, luxb_vid* row_right, luxb_eid* col_left,
                          uint64_t* fq_left, uint64_t* fq_right) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  uint64_t fsize = 0;
  for (int p = 0; p < g->P; ++p) {  // always the reference's split (Graph::Graph, pull_model.inl:108-131)
    if (row_left) row_left[p] = g->ref_rl[p];
    if (row_right) row_right[p] = g->ref_rl[p] + g->ref_np[p] - 1;
    if (col_left) col_left[p] = g->ref_cl[p];
    const uint32_t span = g->ref_np[p] ? g->ref_np[p] - 1 : 0;
    uint64_t bytes = 8 + (uint64_t)(span / 16 + 100) * 4;  // sizeof(FrontierHeader) + mySlots * sizeof(V_ID), push_model.inl:393
    if (fq_left) fq_left[p] = fsize;
    fsize += bytes;
    if (fq_right) fq_right[p] = fsize - 1;
  }
  return g->parts_found;
}

int luxb_work_bounds(const luxb_graph* g, luxb_vid* row_left, luxb_vid* row_right, luxb_eid* col_left) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  for (int p = 0; p < g->P; ++p) {
    if (row_left) row_left[p] = g->rl[p];
    if (row_right) row_right[p] = g->rl[p] + g->np[p] - 1;
    if (col_left) col_left[p] = g->cl[p];
  }
  return use_balanced_split(&g->cfg) ? 1 : 0;
}

// ---- communicator ------------------------------------------------------------------------------------------
int luxb_comm_unique_id(char id[LUXB_UNIQUE_ID_BYTES]) {
  LUXB_ARG(id != nullptr, "id is NULL");
  const char* err = nccl().load();
  if (err) { set_error("%s", err); return LUXB_ERR_COMM; }
  ncclUniqueId uid;
  LUXB_NCCL(nccl().GetUniqueId(&uid));
  memcpy(id, &uid, LUXB_UNIQUE_ID_BYTES);
  return 0;
}

int luxb_comm_init(luxb_graph* g, const char id[LUXB_UNIQUE_ID_BYTES]) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  if (g->P == 1) return 0;
  LUXB_ARG(id != nullptr, "id is NULL");
  LUXB_ARG(g->comm == nullptr, "communicator already initialised");
  const char* err = nccl().load();
  if (err) { set_error("%s", err); return LUXB_ERR_COMM; }
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  ncclUniqueId uid;
  memcpy(&uid, id, LUXB_UNIQUE_ID_BYTES);
  LUXB_NCCL(nccl().CommInitRank(&g->comm, g->P, uid, g->cfg.rank));
  // second stream: the cold half of the PageRank exchange overlaps with the next sweep's panel kernel
  LUXB_CUDA(cudaStreamCreateWithFlags(&g->stream2, cudaStreamNonBlocking));
  LUXB_CUDA(cudaEventCreateWithFlags(&g->ev_pack, cudaEventDisableTiming));
  LUXB_CUDA(cudaEventCreateWithFlags(&g->ev_cold, cudaEventDisableTiming));
  return 0;
}

struct P2PBlob {
  cudaIpcMemHandle_t val[2];  // natural-order replicas (col_filter stores into its peers' replicas)
  cudaIpcMemHandle_t xt[2];   // PageRank: packed transfer arrays
  cudaIpcMemHandle_t fq;      // CC / SSSP: frontier slots of every partition (val[0] = label replica)
  cudaIpcMemHandle_t flags;   // barrier flag words (flag_barrier_kernel)
  int has_val, has_xt, has_fq, has_flags;
};

int luxb_p2p_export(luxb_graph* g, void* blob, size_t* blob_bytes) {
  LUXB_ARG(g && blob_bytes, "NULL argument");
  if (!blob) { *blob_bytes = sizeof(P2PBlob); return 0; }
  LUXB_ARG(*blob_bytes >= sizeof(P2PBlob), "blob too small");
  if (!g->inited) { set_error("luxb_p2p_export: call luxb_init first"); return LUXB_ERR_STATE; }
  P2PBlob b;
  memset(&b, 0, sizeof(b));
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  if (g->cfg.app != LUXB_PAGERANK) {
    for (int k = 0; k < 2; ++k)
      if (g->d_val[k]) LUXB_CUDA(cudaIpcGetMemHandle(&b.val[k], g->d_val[k]));
    b.has_val = 1;
  }
  if (g->d_fq_all) { LUXB_CUDA(cudaIpcGetMemHandle(&b.fq, g->d_fq_all)); b.has_fq = 1; }
  if (g->packed) {
    for (int k = 0; k < 2; ++k) LUXB_CUDA(cudaIpcGetMemHandle(&b.xt[k], g->d_xt[k]));
    b.has_xt = 1;
  }
  if (g->flag_barrier) {
    if (!g->d_flags) {
      LUXB_TRY(dmalloc(&g->d_flags, LUXB_MAX_PARTS));
      LUXB_CUDA(cudaMemset(g->d_flags, 0, sizeof(uint32_t) * LUXB_MAX_PARTS));  // before any peer can learn the handle
      LUXB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g->h_barrier_err), 4, cudaHostAllocMapped));
      *g->h_barrier_err = 0;
    }
    LUXB_CUDA(cudaIpcGetMemHandle(&b.flags, g->d_flags));
    b.has_flags = 1;
  }
  memcpy(blob, &b, sizeof(b));
  *blob_bytes = sizeof(P2PBlob);
  return 0;
}

int luxb_p2p_import(luxb_graph* g, const void* all_blobs, size_t blob_bytes_each) {
  LUXB_ARG(g && all_blobs, "NULL argument");
  LUXB_ARG(blob_bytes_each == sizeof(P2PBlob), "blob size mismatch");
  if (!g->inited) { set_error("luxb_p2p_import: call luxb_init first"); return LUXB_ERR_STATE; }
  if (g->P == 1) return 0;
  LUXB_ARG(g->comm != nullptr, "P2P exchange still needs the communicator for its iteration barrier");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  const P2PBlob* blobs = reinterpret_cast<const P2PBlob*>(all_blobs);
  auto undo = [&]() {  // an import that fails half-way leaves nothing mapped
    for (int p = 0; p < g->P; ++p) {
      if (p == g->cfg.rank) continue;
      for (int k = 0; k < 2; ++k) {
        if (g->peer_val[k][p]) { cudaIpcCloseMemHandle(g->peer_val[k][p]); g->peer_val[k][p] = nullptr; }
        if (g->peer_xt[k][p]) { cudaIpcCloseMemHandle(g->peer_xt[k][p]); g->peer_xt[k][p] = nullptr; }
      }
      if (g->peer_fq[p]) { cudaIpcCloseMemHandle(g->peer_fq[p]); g->peer_fq[p] = nullptr; }
      if (g->peer_flags[p]) { cudaIpcCloseMemHandle(g->peer_flags[p]); g->peer_flags[p] = nullptr; }
    }
  };
  bool all_flags = g->flag_barrier && g->d_flags;
  for (int p = 0; p < g->P; ++p) all_flags = all_flags && (p == g->cfg.rank || blobs[p].has_flags);
  for (int p = 0; p < g->P; ++p) {
    if (p == g->cfg.rank) {
      for (int k = 0; k < 2; ++k) { g->peer_val[k][p] = g->d_val[k]; g->peer_xt[k][p] = g->d_xt[k]; }
      g->peer_fq[p] = g->d_fq_all;
      g->peer_flags[p] = g->d_flags;
      continue;
    }
    if (all_flags) {
      cudaError_t e = cudaIpcOpenMemHandle(&g->peer_flags[p], blobs[p].flags, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle (rank %d's barrier flags): %s", p, cudaGetErrorString(e));
        undo();
        return LUXB_ERR_CUDA;
      }
    }
    if (blobs[p].has_fq && g->d_fq_all) {
      cudaError_t e = cudaIpcOpenMemHandle(&g->peer_fq[p], blobs[p].fq, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle (rank %d's frontier slots): %s", p, cudaGetErrorString(e));
        undo();
        return LUXB_ERR_CUDA;
      }
    }
    for (int k = 0; k < 2; ++k) {
      cudaError_t e = cudaSuccess;
      if (blobs[p].has_val && g->d_val[k]) e = cudaIpcOpenMemHandle(&g->peer_val[k][p], blobs[p].val[k], cudaIpcMemLazyEnablePeerAccess);
      if (e == cudaSuccess && blobs[p].has_xt && g->d_xt[k])
        e = cudaIpcOpenMemHandle(&g->peer_xt[k][p], blobs[p].xt[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("cudaIpcOpenMemHandle (rank %d's buffers): %s", p, cudaGetErrorString(e));
        undo();
        return LUXB_ERR_CUDA;
      }
    }
  }
  g->flag_barrier = all_flags;  // every rank decides alike: the blobs are the same everywhere
  g->p2p_ready = true;
  return 0;
}

int luxb_p2p_disable(luxb_graph* g) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  g->p2p_ready = false;
  return 0;
}

// close every peer buffer mapped into this process
static void p2p_unmap(luxb_graph* g) {
  for (int p = 0; p < g->P; ++p) {
    if (p == g->cfg.rank) continue;
    for (int k = 0; k < 2; ++k) {
      if (g->peer_val[k][p]) { cudaIpcCloseMemHandle(g->peer_val[k][p]); g->peer_val[k][p] = nullptr; }
      if (g->peer_xt[k][p]) { cudaIpcCloseMemHandle(g->peer_xt[k][p]); g->peer_xt[k][p] = nullptr; }
    }
    if (g->peer_fq[p]) { cudaIpcCloseMemHandle(g->peer_fq[p]); g->peer_fq[p] = nullptr; }
    if (g->peer_flags[p]) { cudaIpcCloseMemHandle(g->peer_flags[p]); g->peer_flags[p] = nullptr; }
  }
  g->p2p_ready = false;
}

int luxb_p2p_disconnect(luxb_graph* g) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  if (g->stream) LUXB_CUDA(cudaStreamSynchronize(g->stream));
  if (g->stream2) LUXB_CUDA(cudaStreamSynchronize(g->stream2));  // a cold pull may still be reading the peers' transfer arrays
  p2p_unmap(g);
  return 0;
}

// ---- init ---------------------------------------------------------------------------------------------------
extern "C++" {
// temporaries of a build step: freed on every exit path
struct DevTmp {
  std::vector<void*> ptrs;
  ~DevTmp() { for (void* q : ptrs) cudaFree(q); }
  template <class T>
  int alloc(T** out, uint64_t count) {
    LUXB_TRY(dmalloc(out, count));
    ptrs.push_back(*out);
    return 0;
  }
  void release(void* q) {
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == q) { cudaFree(q); ptrs.erase(ptrs.begin() + i); return; }
  }
  void keep(void* q) {  // ownership moves to the graph
    for (size_t i = 0; i < ptrs.size(); ++i)
      if (ptrs[i] == q) { ptrs.erase(ptrs.begin() + i); return; }
  }
};
}  // extern "C++"
static int build_push_csr(luxb_graph* g) {
  // CSR-by-source over this partition's own edges (init_push_* kernels, components_gpu.cu:550-607):
  // stable radix sort of (src, dst) pairs by src keeps each source's destinations ascending -> deterministic.
  DevTmp tmp;  // temporaries are released on every exit path
  LUXB_TRY(dmalloc(&g->d_out_end, g->nv));
  LUXB_TRY(dmalloc(&g->d_out_dst, g->e_part));
  uint32_t* d_cnt = nullptr;
  LUXB_TRY(tmp.alloc(&d_cnt, g->nv));
  LUXB_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)g->nv * 4, g->stream));
  const int grid = g->num_sms * 8;
  hist_src_kernel<<<grid, 256, 0, g->stream>>>(g->d_src, g->e_part, d_cnt);
  widen_u32_to_u64_kernel<<<grid, 256, 0, g->stream>>>(d_cnt, g->d_out_end, g->nv);
  size_t tmp_bytes = 0;
  LUXB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, g->d_out_end, g->d_out_end, (int)g->nv, g->stream));
  void* d_tmp = nullptr;
  LUXB_TRY(tmp.alloc((char**)&d_tmp, tmp_bytes + 256));
  LUXB_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, g->d_out_end, g->d_out_end, (int)g->nv, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  tmp.release(d_tmp);
  tmp.release(d_cnt);
  if (g->e_part == 0) return 0;
  uint32_t *d_dst = nullptr, *d_keys_out = nullptr;
  LUXB_TRY(tmp.alloc(&d_dst, g->e_part));
  LUXB_TRY(tmp.alloc(&d_keys_out, g->e_part));
  edge_dst_kernel<<<grid, 256, 0, g->stream>>>(g->d_row_end, g->n_part, g->e_part, g->row_left, d_dst);
  LUXB_CUDA(cudaGetLastError());
  int vbits = 1;
  while ((1ull << vbits) < (uint64_t)g->nv) ++vbits;
  tmp_bytes = 0;
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, g->d_src, d_keys_out, d_dst, g->d_out_dst, (long long)g->e_part, 0,
                                            vbits, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_tmp, tmp_bytes + 256));
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, g->d_src, d_keys_out, d_dst, g->d_out_dst, (long long)g->e_part, 0,
                                            vbits, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  return 0;
}

static unsigned char* slot_ptr(unsigned char* base, const luxb_graph* g, int p) { return base + g->slot_off[p]; }

// initial labels + frontier (components_gpu.cu:733-739; sssp_gpu.cu:733-744) — every rank builds all slots locally
static int reset_label_state(luxb_graph* g, bool all_active) {
  const bool cc = g->cfg.app == LUXB_CC;
  uint32_t* lab = reinterpret_cast<uint32_t*>(g->d_val[0]);
  LUXB_CUDA(cudaMemsetAsync(g->d_fq_all, 0, g->fq_total, g->stream));
  for (int p = 0; p < g->P; ++p) {
    FrontierHeader h;
    unsigned char* slot = slot_ptr(g->d_fq_all, g, p);
    if (cc || all_active) {
      h.type = LUXB_DENSE_BITMAP;
      h.num_nodes = g->np[p];
      uint64_t bytes = g->np[p] ? (g->np[p] - 1) / 8 + 1 : 0;  // (R-L)/8 + 1, components_gpu.cu:737
      if (bytes) LUXB_CUDA(cudaMemsetAsync(slot + 8, 0xFF, bytes, g->stream));
    } else {
      h.type = LUXB_SPARSE_QUEUE;
      bool mine = g->np[p] && g->cfg.start_vtx >= g->rl[p] && g->cfg.start_vtx - g->rl[p] < g->np[p];
      h.num_nodes = mine ? 1 : 0;
      if (mine) {
        uint32_t q[1] = {g->cfg.start_vtx};
        uint32_t zero[1] = {0};
        LUXB_CUDA(cudaMemcpyAsync(slot + 8, q, 4, cudaMemcpyHostToDevice, g->stream));
        LUXB_CUDA(cudaMemcpyAsync(slot + 8 + (size_t)g->cap[p] * 4, zero, 4, cudaMemcpyHostToDevice, g->stream));
      }
    }
    LUXB_CUDA(cudaMemcpyAsync(slot, &h, 8, cudaMemcpyHostToDevice, g->stream));
    g->h_hdr[2 * p] = h.type;
    g->h_hdr[2 * p + 1] = h.num_nodes;
  }
  if (!all_active) {
    const int grid = g->num_sms * 8;
    if (cc) iota_kernel<<<grid, 256, 0, g->stream>>>(lab, g->nv);
    else {
      fill_kernel<uint32_t><<<grid, 256, 0, g->stream>>>(lab, g->nv, g->nv);
      uint32_t zero = 0;
      if (g->cfg.start_vtx < g->nv)
        LUXB_CUDA(cudaMemcpyAsync(lab + g->cfg.start_vtx, &zero, 4, cudaMemcpyHostToDevice, g->stream));
    }
    LUXB_CUDA(cudaGetLastError());
  }
  if (g->n_part)
    LUXB_CUDA(cudaMemcpyAsync(g->d_cur, lab + g->row_left, (size_t)g->n_part * 4, cudaMemcpyDeviceToDevice, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  g->stats.last_active = 0;
  for (int p = 0; p < g->P; ++p) g->stats.last_active += g->h_hdr[2 * p + 1];
  g->stats.last_frontier_type = g->h_hdr[2 * g->cfg.rank];
  return 0;
}

// Pin the hot copies in L2: persisting access-policy window on the hot buffer for every kernel of this stream,
// everything else is treated as streaming when it misses.  LUXB_L2_PERSIST=0 disables.
static int set_l2_persisting_window(luxb_graph* g, void* base, size_t bytes) {
  if (const char* env = getenv("LUXB_L2_PERSIST")) if (atoi(env) == 0) return 0;
  int max_persist = 0, max_window = 0;
  LUXB_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, g->cfg.device));
  LUXB_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, g->cfg.device));
  if (max_persist <= 0 || max_window <= 0) return 0;
  if (const char* env = getenv("LUXB_L2_WINDOW_MB")) bytes = std::min<size_t>(bytes, (size_t)(atof(env) * 1e6));  // hottest prefix only
  size_t persist = std::min<size_t>((size_t)max_persist, bytes);
  LUXB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
  cudaStreamAttrValue attr{};
  attr.accessPolicyWindow.base_ptr = base;
  attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)max_window);
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)attr.accessPolicyWindow.num_bytes);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  LUXB_CUDA(cudaStreamSetAttribute(g->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  if (g->cfg.verbose) printf("L2 persisting window: %zu bytes (device max persisting %d, max window %d)\n", persist, max_persist, max_window);
  return 0;
}

// Choose the hot set (largest out-degrees, at most LUXB_HOT_MB megabytes of values, default 64 MB ~ half of L2) and
// rewrite this partition's source ids as indices into the gather space Z = [hot copies in global hotness order | cold]
// (see build.cuh).  cold = the natural-order value array, or — compact_cold, the packed exchange of PageRank on several
// ranks — only the cold vertices that are ever gathered, in id order.
static int build_hot_layout(luxb_graph* g, bool compact_cold) {
  g->hot_n = 0;
  g->packed = false;
  double hot_mb = 64.0;
  if (const char* env = getenv("LUXB_HOT_MB")) hot_mb = atof(env);
  uint64_t h_max = (uint64_t)(hot_mb * 1e6 / 4.0);
  if (h_max == 0 || g->nv < 2 || (uint64_t)g->nv >= 0xFFFFFFFFull - h_max) return 0;
  if (h_max > g->nv) h_max = g->nv;
  const int grid = g->num_sms * 8;
  const uint32_t cap = 4096;
  DevTmp tmp;
  unsigned long long* d_hist = nullptr;
  LUXB_TRY(tmp.alloc(&d_hist, cap + 1));
  LUXB_CUDA(cudaMemsetAsync(d_hist, 0, (cap + 1) * 8, g->stream));
  degree_hist_kernel<<<grid, 256, 0, g->stream>>>(g->d_deg, g->nv, cap, d_hist);
  std::vector<unsigned long long> hist(cap + 1);
  LUXB_CUDA(cudaMemcpyAsync(hist.data(), d_hist, (cap + 1) * 8, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  // smallest tau >= 2 with |{deg >= tau}| <= h_max  (degree-1 vertices are gathered once: packing cannot help them)
  uint64_t above = 0;
  uint32_t tau = cap + 1;
  for (uint32_t d = cap; d >= 2; --d) {
    if (above + hist[d] > h_max) break;
    above += hist[d];
    tau = d;
  }
  if (above == 0 || tau > cap) return 0;
  const uint32_t H = (uint32_t)above;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  uint32_t *d_ids = nullptr, *d_ids2 = nullptr, *d_map = nullptr;
  unsigned int* d_cursor = nullptr;  // [0] cursor, [1 .. P] per-owner counts
  LUXB_TRY(tmp.alloc(&d_keys, H));
  LUXB_TRY(tmp.alloc(&d_keys2, H));
  LUXB_TRY(tmp.alloc(&d_ids, H));
  LUXB_TRY(tmp.alloc(&d_ids2, H));
  LUXB_TRY(dmalloc(&g->d_hot_order, H));
  LUXB_TRY(tmp.alloc(&d_cursor, 1 + LUXB_MAX_PARTS));
  LUXB_CUDA(cudaMemsetAsync(d_cursor, 0, 4 * (1 + LUXB_MAX_PARTS), g->stream));
  PartTable pt{};
  pt.P = g->P;
  for (int p = 0; p < g->P; ++p) { pt.rl[p] = g->rl[p]; pt.np[p] = g->np[p]; }
  hot_select_kernel<<<grid, 256, 0, g->stream>>>(g->d_deg, g->nv, tau, pt, d_cursor, d_cursor + 1, d_keys, d_ids, H);
  LUXB_CUDA(cudaGetLastError());
  // ids arrive in nondeterministic order: sort by (key, id) = two stable passes (id first, then key)
  size_t tb = 0;
  void* d_tmp = nullptr;
  int vbits = 1;
  while ((1ull << vbits) < (uint64_t)g->nv) ++vbits;
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_ids, d_ids2, d_keys, d_keys2, (int)H, 0, vbits, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_ids, d_ids2, d_keys, d_keys2, (int)H, 0, vbits, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  tmp.release(d_tmp);
  tb = 0;
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_keys2, d_keys, d_ids2, g->d_hot_order, (int)H, 0, 32, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys2, d_keys, d_ids2, g->d_hot_order, (int)H, 0, 32, g->stream));
  unsigned int h_cnt[1 + LUXB_MAX_PARTS];
  LUXB_CUDA(cudaMemcpyAsync(h_cnt, d_cursor, sizeof(h_cnt), cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  tmp.release(d_tmp);
  g->hot_off[0] = 0;
  for (int p = 0; p < g->P; ++p) g->hot_off[p + 1] = g->hot_off[p] + h_cnt[1 + p];
  LUXB_TRY(tmp.alloc(&d_map, g->nv));
  if (compact_cold && g->P > 1) {
    // cold-active vertices (0 < deg < tau), ranked in id order; owner p's share is [cold_off[p], cold_off[p+1])
    uint32_t *d_cflag = nullptr, *d_crank = nullptr, *d_okeys = nullptr, *d_okeys2 = nullptr, *d_ranks = nullptr;
    LUXB_TRY(tmp.alloc(&d_cflag, (uint64_t)g->nv + 1));
    LUXB_TRY(tmp.alloc(&d_crank, (uint64_t)g->nv + 1));
    cold_flag_kernel<<<grid, 256, 0, g->stream>>>(g->d_deg, g->nv, tau, d_cflag);
    LUXB_CUDA(cudaMemsetAsync(d_cflag + g->nv, 0, 4, g->stream));
    tb = 0;
    LUXB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cflag, d_crank, (int)g->nv + 1, g->stream));
    LUXB_TRY(tmp.alloc((char**)&d_tmp, tb + 256));
    LUXB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_cflag, d_crank, (int)g->nv + 1, g->stream));
    for (int p = 0; p <= g->P; ++p) {
      const uint32_t at = p < g->P ? std::min(g->rl[p], g->nv) : g->nv;  // empty partitions sit at nv
      LUXB_CUDA(cudaMemcpyAsync(&g->cold_off[p], d_crank + at, 4, cudaMemcpyDeviceToHost, g->stream));
    }
    LUXB_CUDA(cudaStreamSynchronize(g->stream));
    tmp.release(d_tmp);
    g->cold_n = g->cold_off[g->P];
    gather_map_compact_kernel<<<grid, 256, 0, g->stream>>>(d_map, d_crank, g->nv, H);
    // transfer order of the hot values: grouped by owner (stable: hotness order inside a group)
    LUXB_TRY(tmp.alloc(&d_okeys, H));
    LUXB_TRY(tmp.alloc(&d_okeys2, H));
    LUXB_TRY(tmp.alloc(&d_ranks, H));
    LUXB_TRY(dmalloc(&g->d_zperm, H));
    owner_keys_kernel<<<grid, 256, 0, g->stream>>>(g->d_hot_order, H, pt, d_okeys, d_ranks);
    LUXB_CUDA(cudaGetLastError());
    tb = 0;
    LUXB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_okeys, d_okeys2, d_ranks, g->d_zperm, (int)H, 0, 8, g->stream));
    LUXB_TRY(tmp.alloc((char**)&d_tmp, tb + 256));
    LUXB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_okeys, d_okeys2, d_ranks, g->d_zperm, (int)H, 0, 8, g->stream));
    const int me = g->cfg.rank;
    const uint32_t nh_me = g->hot_off[me + 1] - g->hot_off[me], nc_me = g->cold_off[me + 1] - g->cold_off[me];
    LUXB_TRY(dmalloc(&g->d_pack_list, (uint64_t)nh_me + nc_me + 1));
    if (nh_me)
      pack_list_hot_kernel<<<grid_for(nh_me, 256, grid), 256, 0, g->stream>>>(g->d_hot_order, g->d_zperm, g->hot_off[me], nh_me, g->row_left,
                                                                              g->d_pack_list);
    if (g->n_part)
      pack_list_cold_kernel<<<grid, 256, 0, g->stream>>>(d_cflag, d_crank, g->row_left, g->n_part, g->cold_off[me], g->d_pack_list + nh_me);
    LUXB_CUDA(cudaGetLastError());
    LUXB_CUDA(cudaStreamSynchronize(g->stream));
    tmp.release(d_tmp);
    g->packed = true;
  } else {
    gather_map_init_kernel<<<grid, 256, 0, g->stream>>>(d_map, g->nv, H);
  }
  gather_map_hot_kernel<<<grid, 256, 0, g->stream>>>(d_map, g->d_hot_order, H);
  LUXB_TRY(edge_alloc(g, &g->d_src_gather, g->e_part + 8));
  LUXB_CUDA(cudaMemsetAsync(g->d_src_gather, 0, (g->e_part + 8) * 4, g->stream));
  remap_src_kernel<<<grid, 256, 0, g->stream>>>(g->d_src, g->e_part, d_map, g->d_src_gather);
  LUXB_CUDA(cudaGetLastError());
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  g->hot_n = H;
  return 0;
}

static int allgather_slices(luxb_graph* g, void* replica, size_t elem_bytes);
static int build_seg_sweep(luxb_graph* g);
static int pagerank_publish(luxb_graph* g, float* x_new);
static int wait_cold_exchange(luxb_graph* g);

int luxb_init(luxb_graph* g) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  if (g->inited) { set_error("luxb_init called twice"); return LUXB_ERR_STATE; }
  if (g->P > 1 && !g->comm) { set_error("luxb_init: nranks > 1 needs luxb_comm_init first"); return LUXB_ERR_STATE; }
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  const int grid = g->num_sms * 8;
  switch (g->cfg.app) {
    case LUXB_PAGERANK: {
      g->vbytes = 4;
      LUXB_TRY(dmalloc(&g->d_deg, g->nv));
      LUXB_CUDA(cudaMemsetAsync(g->d_deg, 0, (size_t)g->nv * 4, g->stream));
      hist_src_kernel<<<grid, 256, 0, g->stream>>>(g->d_src, g->e_part, g->d_deg);  // pull_scan_task_impl
      LUXB_CUDA(cudaGetLastError());
      if (g->P > 1) LUXB_NCCL(nccl().AllReduce(g->d_deg, g->d_deg, g->nv, ncclUint32, ncclSum, g->comm, g->stream));
      LUXB_TRY(build_hot_layout(g, /*compact_cold=*/true));
      LUXB_TRY(build_seg_sweep(g));
      for (int k = 0; k < 2; ++k) LUXB_TRY(dmalloc((float**)&g->d_val[k], (uint64_t)g->nv + 64));
      pr_init_kernel<<<grid, 256, 0, g->stream>>>(g->d_deg, g->nv, (float*)g->d_val[0]);
      LUXB_CUDA(cudaMemsetAsync(g->d_val[1], 0, (size_t)g->nv * 4, g->stream));
      if (g->hot_n) {
        // + one whole table of slack: the panel kernel always bulk-loads full blocks (panel.cuh)
        LUXB_TRY(dmalloc((float**)&g->d_hot, (uint64_t)g->hot_n + 65536));
        LUXB_CUDA(cudaMemsetAsync(g->d_hot, 0, ((size_t)g->hot_n + 65536) * 4, g->stream));
        LUXB_TRY(set_l2_persisting_window(g, g->d_hot, (size_t)g->hot_n * 4));
      }
      if (g->packed) {
        g->xt_hot_chunk = ((((uint64_t)g->hot_n + g->P - 1) / g->P) + 31) & ~31ull;
        g->xt_cold_chunk = ((((uint64_t)g->cold_n + g->P - 1) / g->P) + 31) & ~31ull;
        const uint64_t xt_len = (g->xt_hot_chunk + g->xt_cold_chunk) * g->P;
        for (int k = 0; k < 2; ++k) {
          LUXB_TRY(dmalloc(&g->d_xt[k], xt_len));
          LUXB_CUDA(cudaMemsetAsync(g->d_xt[k], 0, xt_len * 4, g->stream));
        }
      }
      LUXB_CUDA(cudaGetLastError());
      // every rank holds the complete x0: publish it (packs this rank's share, exchanges, fills the hot copies)
      LUXB_TRY(pagerank_publish(g, (float*)g->d_val[0]));
      g->replica_stale = false;
      break;
    }
    case LUXB_COLFILTER: {
      g->vbytes = 4 * kCfK;
      for (int k = 0; k < 2; ++k) LUXB_TRY(dmalloc((float**)&g->d_val[k], (uint64_t)g->nv * kCfK));
      cf_init_kernel<<<grid, 256, 0, g->stream>>>((float*)g->d_val[0], (uint64_t)g->nv * kCfK);
      cf_init_kernel<<<grid, 256, 0, g->stream>>>((float*)g->d_val[1], (uint64_t)g->nv * kCfK);
      // chunk table
      uint32_t* d_cnt = nullptr;
      LUXB_TRY(dmalloc(&d_cnt, (uint64_t)g->n_part + 1));
      LUXB_TRY(dmalloc(&g->d_chunk_first, (uint64_t)g->n_part + 2));
      LUXB_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)g->n_part + 1) * 4, g->stream));
      cf_chunk_count_kernel<<<grid, 256, 0, g->stream>>>(g->d_row_end, g->n_part, d_cnt);
      size_t tmp_bytes = 0;
      LUXB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, g->d_chunk_first, (int)g->n_part + 1, g->stream));
      void* d_tmp = nullptr;
      LUXB_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 256));
      LUXB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, g->d_chunk_first, (int)g->n_part + 1, g->stream));
      LUXB_CUDA(cudaMemcpyAsync(&g->n_chunks, g->d_chunk_first + g->n_part, 4, cudaMemcpyDeviceToHost, g->stream));
      LUXB_CUDA(cudaStreamSynchronize(g->stream));
      LUXB_CUDA(cudaFree(d_tmp));
      LUXB_CUDA(cudaFree(d_cnt));
      LUXB_TRY(dmalloc(&g->d_chunk_vtx, g->n_chunks));
      LUXB_TRY(dmalloc(&g->d_partial, (uint64_t)g->n_chunks * kCfK));
      cf_chunk_fill_kernel<<<grid, 256, 0, g->stream>>>(g->d_chunk_first, g->n_part, g->d_chunk_vtx);
      LUXB_CUDA(cudaGetLastError());
      break;
    }
    case LUXB_CC:
    case LUXB_SSSP: {
      g->vbytes = 4;
      LUXB_ARG(g->cfg.app != LUXB_SSSP || g->cfg.start_vtx < g->nv, "start vertex %u >= nv", g->cfg.start_vtx);
      LUXB_TRY(dmalloc((uint32_t**)&g->d_val[0], g->nv));
      LUXB_TRY(dmalloc(&g->d_cur, g->n_part));
      LUXB_TRY(build_push_csr(g));
      {  // hot-packed label copies for the pull sweeps (same layout as PageRank; refreshed before every pull sweep)
        LUXB_TRY(dmalloc(&g->d_deg, g->nv));
        LUXB_CUDA(cudaMemsetAsync(g->d_deg, 0, (size_t)g->nv * 4, g->stream));
        hist_src_kernel<<<grid, 256, 0, g->stream>>>(g->d_src, g->e_part, g->d_deg);
        LUXB_CUDA(cudaGetLastError());
        if (g->P > 1) LUXB_NCCL(nccl().AllReduce(g->d_deg, g->d_deg, g->nv, ncclUint32, ncclSum, g->comm, g->stream));
        LUXB_TRY(build_hot_layout(g, /*compact_cold=*/false));
        LUXB_TRY(build_seg_sweep(g));
        if (g->hot_n) {
          LUXB_TRY(dmalloc((uint32_t**)&g->d_hot, (uint64_t)g->hot_n + 65536));  // + one table of slack (panel.cuh)
          LUXB_CUDA(cudaMemsetAsync(g->d_hot, 0, ((size_t)g->hot_n + 65536) * 4, g->stream));
          LUXB_TRY(set_l2_persisting_window(g, g->d_hot, (size_t)g->hot_n * 4));
        }
      }
      g->big_capacity = (uint32_t)std::min<uint64_t>(g->e_part / kPushBigDegree + 1024, 0x7FFFFFFFull);
      LUXB_TRY(dmalloc((PushArgs::BigSeg**)&g->d_big_list, g->big_capacity));
      LUXB_TRY(dmalloc(&g->d_fq_all, g->fq_total));
      LUXB_TRY(dmalloc(&g->d_fq_new, g->slot_bytes[g->cfg.rank]));
      LUXB_TRY(dmalloc(&g->d_fq_tmp, g->slot_bytes[g->cfg.rank]));
      LUXB_TRY(dmalloc(&g->d_hdr_all, 2 * LUXB_MAX_PARTS));
      LUXB_CUDA(cudaMallocHost(&g->h_hdr, 2 * LUXB_MAX_PARTS * 4));
      LUXB_CUDA(cudaMallocHost(&g->h_scratch, 64));
      LUXB_TRY(reset_label_state(g, false));
      if (g->P > 1) {
        // warm the communicator with the collectives the hot loop uses (NCCL sets up channels lazily, ~1 s for the
        // first large grouped broadcast at 8 ranks): re-broadcasting the identical initial labels is a no-op
        LUXB_NCCL(nccl().AllGather(g->d_fq_new, g->d_hdr_all, 8, ncclUint8, g->comm, g->stream));
        LUXB_TRY(allgather_slices(g, g->d_val[0], 4));
        LUXB_CUDA(cudaStreamSynchronize(g->stream));
      }
      break;
    }
  }
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  g->cur = 0;
  g->inited = true;
  return 0;
}

// ---- exchange helpers ---------------------------------------------------------------------------------------
// all-gather of unequal slices = one grouped set of in-place broadcasts (root p sends partition p's slice)
static int allgather_slices(luxb_graph* g, void* replica, size_t elem_bytes) {
  if (g->P == 1) return 0;
  LUXB_NCCL(nccl().GroupStart());
  for (int p = 0; p < g->P; ++p) {
    if (g->np[p] == 0) continue;
    char* ptr = reinterpret_cast<char*>(replica) + (size_t)g->rl[p] * elem_bytes;
    LUXB_NCCL(nccl().Broadcast(ptr, ptr, (size_t)g->np[p] * elem_bytes, ncclUint8, p, g->comm, g->stream));
  }
  LUXB_NCCL(nccl().GroupEnd());
  return 0;
}

// iteration barrier of the P2P exchange: a 4-byte all-reduce enqueued after the compute kernels; when it
// completes on a rank, every peer's kernels (and therefore their stores into this rank's replica) are done.
static int p2p_barrier(luxb_graph* g) {
  if (g->P == 1) return 0;
  // the flag kernel is used where it has been validated on hardware (PageRank exchange, 4 GPUs, bit-identical values and the
  // same iteration time as the all-reduce: profiles/r02_trace_exchange_n4.txt); LUXB_BARRIER=flag extends it to the other apps
  if (g->flag_barrier && g->p2p_ready && (g->cfg.app == LUXB_PAGERANK || g->flag_barrier_all)) {
    FlagBarrierArgs a{};
    for (int p = 0; p < g->P; ++p) a.peer[p] = reinterpret_cast<uint32_t*>(g->peer_flags[p]);
    a.mine = g->d_flags;
    a.err = g->h_barrier_err;
    a.P = g->P;
    a.me = g->cfg.rank;
    a.epoch = ++g->barrier_epoch;
    a.timeout_ns = g->barrier_timeout_ns;
    flag_barrier_kernel<<<1, LUXB_MAX_PARTS, 0, g->stream>>>(a);
    LUXB_CUDA(cudaGetLastError());
    g->stats.kernel_launches++;
    return 0;
  }
  if (!g->d_sync) LUXB_TRY(dmalloc(&g->d_sync, 4));
  LUXB_NCCL(nccl().AllReduce(g->d_sync, g->d_sync, 1, ncclUint32, ncclSum, g->comm, g->stream));
  return 0;
}

extern "C++" {
// the partition's own CSC as a layout view (it owns nothing but the fused fix-up's chain state)
static PullLayout& base_layout(luxb_graph* g, const uint32_t* src_idx) {
  PullLayout& L = g->base_view;
  L.d_row_end = g->d_row_end;
  L.d_row_end32 = g->d_row_end32;
  L.d_src = const_cast<uint32_t*>(src_idx);
  L.d_tile_v = g->d_tile_v;
  L.n_vtx = g->n_part;
  L.e_cnt = g->e_part;
  L.n_tiles = g->n_tiles;
  L.d_head = g->d_head;
  L.d_tail = g->d_tail;
  L.d_carry = g->d_carry;
  L.d_carry_flag = g->d_carry_flag;
  L.d_block_agg = g->d_block_agg;
  L.d_block_flag = g->d_block_flag;
  L.n_fix_blocks = g->n_fix_blocks;
  return L;
}

template <class Prog, class Shape>
static int launch_pull_shape(luxb_graph* g, const PullArgs<Prog>& a) {
  auto kern = pull_tile_kernel<Prog, Shape>;
  const int want = g->pull_ctas;
  // carve out exactly `want` CTAs' worth of shared memory; the rest of the 256 KB unified array stays L1.
  // (function attributes are per device and idempotent: set on every launch, no process-global cache)
  LUXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape::kSmemBytes));
  int carve_pct = (int)std::min<size_t>(100, (want * (Shape::kSmemBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
  LUXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve_pct));
  const uint32_t n_super = (a.n_tiles + Shape::kWarps - 1) / Shape::kWarps;
  uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)g->num_sms * want, n_super);
  kern<<<grid, Shape::kThreads, Shape::kSmemBytes, g->stream>>>(a);
  LUXB_CUDA(cudaGetLastError());
  return 0;
}

static int kt_begin(luxb_graph* g) {
  if (!g->kernel_timing) return 0;
  if (g->kt_used + 2 > g->kt_events.size()) {
    cudaEvent_t e0, e1;
    LUXB_CUDA(cudaEventCreate(&e0));
    LUXB_CUDA(cudaEventCreate(&e1));
    g->kt_events.push_back(e0);
    g->kt_events.push_back(e1);
  }
  LUXB_CUDA(cudaEventRecord(g->kt_events[g->kt_used], g->stream));
  return 0;
}
static int kt_end(luxb_graph* g) {
  if (!g->kernel_timing) return 0;
  LUXB_CUDA(cudaEventRecord(g->kt_events[g->kt_used + 1], g->stream));
  g->kt_used += 2;
  return 0;
}

template <class Prog>
static void fill_fixup_args(PullArgs<Prog>& a, const PullLayout& L) {
  a.tile_v = L.d_tile_v;
  a.n_tiles = L.n_tiles;
  a.head_partial = reinterpret_cast<typename Prog::Acc*>(L.d_head);
  a.tail_partial = reinterpret_cast<typename Prog::Acc*>(L.d_tail);
  a.carry = reinterpret_cast<typename Prog::Wide*>(L.d_carry);
  a.carry_flag = L.d_carry_flag;
  a.block_agg = reinterpret_cast<typename Prog::Wide*>(L.d_block_agg);
  a.block_flag = L.d_block_flag;
}

// L is the layout's own descriptor (the fused fix-up keeps its launch epoch there)
template <class Prog>
static int launch_fixup(luxb_graph* g, const PullArgs<Prog>& a, PullLayout& L) {
  if (L.n_tiles <= 1) return 0;
  if (g->fused_fixup) {
    if (!L.d_chain) {
      LUXB_TRY(dmalloc(&L.d_chain, 4ull * L.n_fix_blocks + 4));  // values, status words, [4 n] = ticket counter
      LUXB_CUDA(cudaMemsetAsync(L.d_chain, 0, (4ull * L.n_fix_blocks + 4) * 8, g->stream));
      L.chain_epoch = 0;
    }
    FixupChain<Prog> ch;
    ch.value = L.d_chain;
    ch.status = L.d_chain + 2ull * L.n_fix_blocks;
    ch.ticket = L.d_chain + 4ull * L.n_fix_blocks;
    ch.n_blocks = L.n_fix_blocks;
    ch.epoch = ++L.chain_epoch;
    if (L.chain_epoch >= 0x3FFFFFF0u) {  // 30-bit epochs: start over with a clean status array
      LUXB_CUDA(cudaMemsetAsync(L.d_chain, 0, (4ull * L.n_fix_blocks + 4) * 8, g->stream));
      L.chain_epoch = ch.epoch = 1;
    }
    pull_fixup_fused_kernel<Prog><<<L.n_fix_blocks, kFixBlock, 0, g->stream>>>(a, ch);
    LUXB_CUDA(cudaGetLastError());
    g->stats.kernel_launches++;
    return 0;
  }
  pull_fixup_scan_kernel<Prog><<<L.n_fix_blocks, kFixBlock, 0, g->stream>>>(a);
  pull_fixup_blocks_kernel<Prog><<<1, 1024, 0, g->stream>>>(a, L.n_fix_blocks);
  pull_fixup_apply_kernel<Prog><<<L.n_fix_blocks, kFixBlock, 0, g->stream>>>(a);
  LUXB_CUDA(cudaGetLastError());
  g->stats.kernel_launches += 3;
  return 0;
}

// one pull sweep over layout L.  hub_bits != nullptr: those vertices get their raw sum (panel.cuh).
template <class Prog>
static int launch_pull(luxb_graph* g, PullLayout& L, const typename Prog::Vertex* x_nat, const typename Prog::Vertex* x_cold,
                       const typename Prog::Vertex* x_hot, uint32_t hot_n, typename Prog::Vertex* out_local,
                       const typename Prog::Params& prm, const uint32_t* hub_bits = nullptr, bool timed = true) {
  if (L.n_tiles == 0) return 0;
  PullArgs<Prog> a{};
  a.row_end = L.d_row_end;
  a.row_end32 = L.d_row_end32;
  a.src = reinterpret_cast<const uint32_t*>(L.d_src);
  a.x_nat = x_nat;
  a.n_part = L.n_vtx;
  a.e_part = L.e_cnt;
  a.row_left = g->row_left;
  a.x_old = x_cold;  // gather ids >= hot_n index this array at (id - hot_n)
  a.x_hot = x_hot;
  a.hot_n = hot_n;
  a.out = out_local;
  fill_fixup_args(a, L);
  a.tile_counter = reinterpret_cast<uint32_t*>(g->d_counters + 2);
  LUXB_CUDA(cudaMemsetAsync(a.tile_counter, 0, 4, g->stream));
  a.prm = prm;
  a.hub_bits = hub_bits;
  a.raw_out = 0;
  if (timed) LUXB_TRY(kt_begin(g));
  switch (g->pull_shape) {
#define LUXB_CASE_SHAPE(id, ipt, warps, stages) \
    case id: LUXB_TRY((launch_pull_shape<Prog, PullShape##id>(g, a))); break;
    LUXB_PULL_SHAPES(LUXB_CASE_SHAPE)
    default: set_error("bad pull shape"); return LUXB_ERR_STATE;
  }
  if (timed) LUXB_TRY(kt_end(g));
  g->stats.kernel_launches++;
  pt_mark(g, 0);
  LUXB_TRY(launch_fixup(g, a, L));
  pt_mark(g, 1);
  return 0;
}
}  // extern "C++"

// ---- flagged segmented-scan sweep (seg.cuh) and its source-blocked variant (panel.cuh) ------------------------------
// shapes <consumer warps, ring stages, rounds of 256 edges per warp piece>
#define LUXB_SEG_MAIN_SHAPES(X) X(0, 8, 2, 2) X(1, 8, 3, 1) X(2, 12, 2, 1) X(3, 8, 2, 4) X(4, 16, 2, 1) X(5, 8, 4, 1) X(6, 8, 2, 1) X(7, 6, 2, 2)
#define LUXB_DECL_MSHAPE(id, warps, stages, rounds) using SegMain##id = SegShape<warps, stages, rounds, false, 0>;
LUXB_SEG_MAIN_SHAPES(LUXB_DECL_MSHAPE)
// panel shapes: + shared-memory table capacity (values, <= 32768: 15-bit offsets); one CTA per SM
// (..., edges per lane and round)
#define LUXB_SEG_PANEL_SHAPES(X) \
  X(0, 31, 2, 2, 32768, 8) X(1, 24, 2, 1, 32768, 16) X(2, 16, 2, 2, 32768, 16) X(3, 24, 2, 2, 32768, 8) X(4, 20, 2, 1, 32768, 16) X(5, 16, 2, 4, 32768, 8)
#define LUXB_DECL_PSHAPE(id, warps, stages, rounds, tab, v) using SegPanel##id = SegShape<warps, stages, rounds, true, tab, v>;
LUXB_SEG_PANEL_SHAPES(LUXB_DECL_PSHAPE)
struct SegShapeInfo { int piece, stage_edges, tab; };
#define LUXB_MSHAPE_INFO(id, warps, stages, rounds) {SegMain##id::kPiece, SegMain##id::kStageEdges, 0},
#define LUXB_PSHAPE_INFO(id, warps, stages, rounds, tab, v) {SegPanel##id::kPiece, SegPanel##id::kStageEdges, SegPanel##id::kTab},
static const SegShapeInfo kSegMainInfo[] = {LUXB_SEG_MAIN_SHAPES(LUXB_MSHAPE_INFO)};
static const SegShapeInfo kSegPanelInfo[] = {LUXB_SEG_PANEL_SHAPES(LUXB_PSHAPE_INFO)};
static const int kNumSegMain = sizeof(kSegMainInfo) / sizeof(SegShapeInfo);
static const int kNumSegPanel = sizeof(kSegPanelInfo) / sizeof(SegShapeInfo);

static void free_layout(PullLayout& L) {
  void* ptrs[] = {L.d_row_end, L.d_row_end32, L.d_src, L.d_tile_v, L.d_head, L.d_tail, L.d_carry, L.d_carry_flag, L.d_block_agg, L.d_block_flag,
                  L.d_close, L.d_empty, L.d_empty_hub, L.d_chain};
  for (void* q : ptrs) if (q) cudaFree(q);
  L = PullLayout();
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

extern "C++" {
// Build the flagged stream of a CSC (seg.cuh).  row_end: inclusive end offsets (u64) of n_vtx "vertices" whose edges,
// in CSC order, carry the gather ids `ids`; blk: the vertex / edge ranges of the blocks (one block = plain stream;
// wbase / hshift are filled here).  Every block is padded with 1 .. stage_edges head-flagged dummy words to a whole
// number of stages.  On return L holds words, close list, tile_v (heads before each piece), fix-up scratch and, if
// want_empty, the list of vertices without edges.  super_end (optional) receives the first stage after each block.
template <class Word, class In>
static int build_seg_stream(luxb_graph* g, PullLayout& L, const uint64_t* d_row_end, uint32_t n_vtx, const In* d_ids, uint64_t e_cnt,
                            StreamBlocks& blk, uint32_t stage_edges, uint32_t piece, uint32_t vtx_offset, bool want_empty,
                            const uint32_t* hub_bits, uint32_t* super_end) {
  const int grid = g->num_sms * 8;
  DevTmp tmp;
  L = PullLayout();
  L.n_vtx = n_vtx;
  L.e_cnt = e_cnt;
  uint64_t words = 0, pads_total = 0;
  for (uint32_t b = 0; b < blk.n_blocks; ++b) {
    const uint64_t eb = blk.ebase[b + 1] - blk.ebase[b];
    const uint64_t pad = stage_edges - eb % stage_edges;  // 1 .. stage_edges: the first pad closes the block's last vertex
    blk.wbase[b] = words;
    blk.hshift[b] = pads_total;
    words += eb + pad;
    pads_total += pad;
    if (super_end) super_end[b] = (uint32_t)(words / stage_edges);
  }
  blk.wbase[blk.n_blocks] = words;
  blk.hshift[blk.n_blocks] = pads_total;
  LUXB_ARG(words / piece < 0xFFFFFFF0ull && words / stage_edges < 0xFFFFFFF0ull, "stream too large");
  L.n_words = words;
  L.n_stages = (uint32_t)(words / stage_edges);
  L.n_tiles = (uint32_t)(words / piece);
  // 1. non-empty vertices and their rank
  uint32_t *d_flag = nullptr, *d_rank = nullptr;
  LUXB_TRY(tmp.alloc(&d_flag, (uint64_t)n_vtx + 1));
  LUXB_TRY(tmp.alloc(&d_rank, (uint64_t)n_vtx + 1));
  nonempty_flag_kernel<<<grid, 256, 0, g->stream>>>(d_row_end, n_vtx, d_flag);
  LUXB_CUDA(cudaMemsetAsync(d_flag + n_vtx, 0, 4, g->stream));
  size_t tb = 0;
  void* d_tmp = nullptr;
  LUXB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_rank, (int)n_vtx + 1, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_rank, (int)n_vtx + 1, g->stream));
  uint32_t n_seg = 0;
  LUXB_CUDA(cudaMemcpyAsync(&n_seg, d_rank + n_vtx, 4, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  tmp.release(d_tmp);
  // 2. words: ids, then pads, then head flags; close list (entry j + 1 = owner of head j, dummies for the pads)
  Word* d_words = nullptr;
  LUXB_TRY(dmalloc(&d_words, words + 64));
  L.d_src = d_words;
  if (e_cnt) stream_copy_kernel<Word, In><<<grid, 256, 0, g->stream>>>(d_ids, e_cnt, blk, d_words);
  for (uint32_t b = 0; b < blk.n_blocks; ++b) {
    const uint64_t from = blk.wbase[b] + (blk.ebase[b + 1] - blk.ebase[b]);
    stream_pad_kernel<Word><<<grid_for(blk.wbase[b + 1] - from, 256, grid), 256, 0, g->stream>>>(d_words, from, blk.wbase[b + 1]);
  }
  LUXB_CUDA(cudaGetLastError());
  const uint64_t n_heads = (uint64_t)n_seg + pads_total;
  LUXB_ARG(n_heads < 0xFFFFFFF0ull, "too many segments");
  LUXB_TRY(dmalloc(&L.d_close, n_heads + 2));
  LUXB_CUDA(cudaMemsetAsync(L.d_close, 0xFF, (n_heads + 2) * 4, g->stream));
  stream_heads_kernel<Word><<<grid, 256, 0, g->stream>>>(d_row_end, n_vtx, d_flag, d_rank, blk, d_words, L.d_close, vtx_offset);
  LUXB_CUDA(cudaGetLastError());
  // 3. heads before each piece
  uint32_t* d_cnt = nullptr;
  LUXB_TRY(tmp.alloc(&d_cnt, (uint64_t)L.n_tiles + 2));
  LUXB_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)L.n_tiles + 2) * 4, g->stream));
  piece_heads_kernel<Word><<<grid, 256, 0, g->stream>>>(d_words, L.n_tiles, piece, d_cnt);
  LUXB_CUDA(cudaGetLastError());
  LUXB_TRY(dmalloc(&L.d_tile_v, (uint64_t)L.n_tiles + 2));
  tb = 0;
  LUXB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt, L.d_tile_v, (int)L.n_tiles + 1, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_cnt, L.d_tile_v, (int)L.n_tiles + 1, g->stream));
  uint32_t heads_counted = 0;
  LUXB_CUDA(cudaMemcpyAsync(&heads_counted, L.d_tile_v + L.n_tiles, 4, cudaMemcpyDeviceToHost, g->stream));
  // 4. vertices without edges (hubs among them apart)
  unsigned int* d_cur2 = nullptr;
  if (want_empty) {
    const uint32_t n_e = n_vtx - n_seg;
    LUXB_TRY(dmalloc(&L.d_empty, (uint64_t)n_e + 1));
    LUXB_TRY(dmalloc(&L.d_empty_hub, (uint64_t)n_e + 1));
    LUXB_TRY(tmp.alloc(&d_cur2, 2));
    LUXB_CUDA(cudaMemsetAsync(d_cur2, 0, 8, g->stream));
    empty_split_kernel<<<grid, 256, 0, g->stream>>>(d_flag, n_vtx, hub_bits, L.d_empty, L.d_empty_hub, d_cur2);
    LUXB_CUDA(cudaGetLastError());
    unsigned int h_cur2[2] = {0, 0};
    LUXB_CUDA(cudaMemcpyAsync(h_cur2, d_cur2, 8, cudaMemcpyDeviceToHost, g->stream));
    LUXB_CUDA(cudaStreamSynchronize(g->stream));
    L.n_empty = h_cur2[0];
    L.n_empty_hub = h_cur2[1];
  }
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  if ((uint64_t)heads_counted != n_heads) {
    set_error("flagged stream: %u heads counted, %llu expected", heads_counted, (unsigned long long)n_heads);
    return LUXB_ERR_STATE;
  }
  // 5. fix-up scratch
  LUXB_TRY(dmalloc((uint32_t**)&L.d_head, (uint64_t)L.n_tiles + 1));
  LUXB_TRY(dmalloc((uint32_t**)&L.d_tail, (uint64_t)L.n_tiles + 1));
  L.n_fix_blocks = (L.n_tiles + kFixBlock - 1) / kFixBlock;
  LUXB_TRY(dmalloc((uint64_t**)&L.d_carry, (uint64_t)L.n_tiles + 1));
  LUXB_TRY(dmalloc(&L.d_carry_flag, (uint64_t)L.n_tiles + 1));
  LUXB_TRY(dmalloc((uint64_t**)&L.d_block_agg, (uint64_t)L.n_fix_blocks + 1));
  LUXB_TRY(dmalloc(&L.d_block_flag, (uint64_t)L.n_fix_blocks + 1));
  return 0;
}
}  // extern "C++"

// the whole partition as one flagged stream (PageRank without the source-blocked split)
static int build_plain_seg_layout(luxb_graph* g) {
  const SegShapeInfo shp = kSegMainInfo[g->seg_main_shape];
  StreamBlocks blk{};
  blk.n_blocks = 1;
  blk.vfirst[0] = 0; blk.vfirst[1] = g->n_part;
  blk.ebase[0] = 0; blk.ebase[1] = g->e_part;
  return build_seg_stream<uint32_t, uint32_t>(g, g->sb_main, g->d_row_end, g->n_part, g->hot_n ? g->d_src_gather : g->d_src, g->e_part, blk,
                                              (uint32_t)shp.stage_edges, (uint32_t)shp.piece, 0, true, nullptr, nullptr);
}

// Split this partition's (hot-packed) CSC into the panel (hot source block x hub destination, 15-bit offsets, gathered
// from shared memory) and the main stream (everything else, gathered through L1).
// LUXB_SB = 0 off / 1 force / unset: automatic (on when the panel would take at least a fifth of a large partition).
// Tuning: LUXB_SB_BS (values per block), LUXB_SB_BLOCKS (max blocks), LUXB_SB_MIN_INDEG (hub threshold).
static int build_panel_layout(luxb_graph* g) {
  g->sb_on = false;
  const int mode = env_int("LUXB_SB", -1);
  if (mode == 0 || g->hot_n == 0 || g->e_part == 0 || g->e_part >= 0xFFFFFFFFull || g->cfg.zero_copy_edges) return 0;
  if (mode < 0 && g->e_part < (1ull << 24)) return 0;
  const SegShapeInfo shp = kSegPanelInfo[g->seg_panel_shape];
  uint32_t bs = (uint32_t)std::max(4, env_int("LUXB_SB_BS", shp.tab));
  bs = std::min<uint32_t>(bs & ~3u, (uint32_t)shp.tab);
  const uint32_t nb_max = (uint32_t)std::min(std::max(env_int("LUXB_SB_BLOCKS", 48), 1), kPanelMaxBlocks);
  const uint32_t n_src = (uint32_t)std::min<uint64_t>(g->hot_n, (uint64_t)nb_max * bs);
  const uint32_t NB = (n_src + bs - 1) / bs;
  const uint32_t min_indeg = (uint32_t)std::max(env_int("LUXB_SB_MIN_INDEG", 64), 1);
  const int grid = g->num_sms * 8;
  DevTmp tmp;

  // 1. hub destinations
  uint32_t *d_flag = nullptr, *d_hub_idx = nullptr;
  LUXB_TRY(tmp.alloc(&d_flag, (uint64_t)g->n_part + 1));
  LUXB_TRY(tmp.alloc(&d_hub_idx, (uint64_t)g->n_part + 1));
  hub_flag_kernel<<<grid, 256, 0, g->stream>>>(g->d_row_end, g->n_part, min_indeg, d_flag);
  LUXB_CUDA(cudaGetLastError());
  size_t tb = 0;
  void* d_scan_tmp = nullptr;
  LUXB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_hub_idx, (int)g->n_part, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_scan_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp, tb, d_flag, d_hub_idx, (int)g->n_part, g->stream));
  uint32_t last_idx = 0, last_flag = 0;
  LUXB_CUDA(cudaMemcpyAsync(&last_idx, d_hub_idx + g->n_part - 1, 4, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaMemcpyAsync(&last_flag, d_flag + g->n_part - 1, 4, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  tmp.release(d_scan_tmp);
  const uint32_t Nh = last_idx + last_flag;
  if (Nh == 0 || (uint64_t)Nh * NB >= 0x7FFFFFF0ull) return 0;
  uint32_t *d_hub_vtx = nullptr, *d_hub_bits = nullptr, *d_cov = nullptr;
  LUXB_TRY(tmp.alloc(&d_hub_vtx, Nh));
  LUXB_TRY(tmp.alloc(&d_hub_bits, ((uint64_t)g->n_part + 31) / 32 + 1));
  LUXB_TRY(tmp.alloc(&d_cov, Nh));
  hub_list_kernel<<<grid, 256, 0, g->stream>>>(d_flag, d_hub_idx, g->n_part, d_hub_vtx, d_hub_bits);
  LUXB_CUDA(cudaGetLastError());

  // 2. edge keys: block of the source for (hot source, hub destination) edges, 255 for the rest; 3. stable sort
  uint8_t *d_key = nullptr, *d_key2 = nullptr;
  uint64_t *d_pay = nullptr, *d_pay2 = nullptr;
  LUXB_TRY(tmp.alloc(&d_key, g->e_part));
  LUXB_TRY(tmp.alloc(&d_key2, g->e_part));
  LUXB_TRY(tmp.alloc(&d_pay, g->e_part));
  LUXB_TRY(tmp.alloc(&d_pay2, g->e_part));
  edge_iota_kernel<<<grid, 256, 0, g->stream>>>(d_pay, d_key, g->e_part);
  hub_key_kernel<<<grid, 256, 0, g->stream>>>(g->d_row_end, g->d_src_gather, d_hub_vtx, Nh, n_src, bs, d_key, d_pay, d_cov);
  LUXB_CUDA(cudaGetLastError());
  tb = 0;
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_key, d_key2, d_pay, d_pay2, (long long)g->e_part, 0, 8, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_scan_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceRadixSort::SortPairs(d_scan_tmp, tb, d_key, d_key2, d_pay, d_pay2, (long long)g->e_part, 0, 8, g->stream));
  unsigned long long* d_hist = nullptr;
  LUXB_TRY(tmp.alloc(&d_hist, 256));
  LUXB_CUDA(cudaMemsetAsync(d_hist, 0, 256 * 8, g->stream));
  key_hist_kernel<<<grid, 256, 0, g->stream>>>(d_key2, g->e_part, d_hist);
  LUXB_CUDA(cudaGetLastError());
  unsigned long long hist[256];
  LUXB_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof(hist), cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  tmp.release(d_scan_tmp);
  tmp.release(d_key);
  tmp.release(d_pay);
  uint64_t e_cov = 0;
  for (uint32_t b = 0; b < NB; ++b) e_cov += hist[b];
  const uint64_t e_main = hist[255];
  if (e_cov + e_main != g->e_part) { set_error("panel split lost edges (%llu + %llu != %llu)", (unsigned long long)e_cov,
                                               (unsigned long long)e_main, (unsigned long long)g->e_part); return LUXB_ERR_STATE; }
  if (e_cov == 0 || (mode < 0 && e_cov < g->e_part / 5)) return 0;

  // 4. panel CSC over virtual vertices (block b, hub h) -> index b * Nh + h: offsets + per-vertex in-degree
  const uint32_t NV = Nh * NB;
  StreamBlocks pblk{};
  pblk.n_blocks = NB;
  for (uint32_t b = 0; b <= NB; ++b) {
    g->sb_pb.vbase[b] = b * Nh;
    pblk.vfirst[b] = b * Nh;
  }
  pblk.ebase[0] = 0;
  for (uint32_t b = 0; b < NB; ++b) pblk.ebase[b + 1] = pblk.ebase[b] + hist[b];
  uint16_t* d_src16 = nullptr;
  uint32_t* d_vcount = nullptr;
  uint64_t* d_vrow = nullptr;
  LUXB_TRY(tmp.alloc(&d_src16, e_cov + 32));
  LUXB_TRY(tmp.alloc(&d_vcount, (uint64_t)NV + 1));
  LUXB_CUDA(cudaMemsetAsync(d_vcount, 0, ((size_t)NV + 1) * 4, g->stream));
  panel_fill_kernel<<<grid, 256, 0, g->stream>>>(d_key2, d_pay2, e_cov, g->d_src_gather, bs, g->sb_pb, d_src16, d_vcount);
  LUXB_CUDA(cudaGetLastError());
  LUXB_TRY(tmp.alloc(&d_vrow, (uint64_t)NV + 4));
  widen_u32_to_u64_kernel<<<grid, 256, 0, g->stream>>>(d_vcount, d_vrow, NV);
  tb = 0;
  LUXB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, d_vrow, d_vrow, (int)NV, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_scan_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceScan::InclusiveSum(d_scan_tmp, tb, d_vrow, d_vrow, (int)NV, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  tmp.release(d_scan_tmp);
  tmp.release(d_vcount);
  LUXB_TRY((build_seg_stream<uint16_t, uint16_t>(g, g->sb_panel, d_vrow, NV, d_src16, e_cov, pblk, (uint32_t)shp.stage_edges,
                                                 (uint32_t)shp.piece, 0, false, nullptr, g->sb_super_end)));
  tmp.release(d_vrow);
  tmp.release(d_src16);

  // 5. main stream: what is left, in the original (dst, src) order
  uint32_t* d_main_src = nullptr;
  uint64_t* d_main_row = nullptr;
  LUXB_TRY(tmp.alloc(&d_main_src, e_main + 8));
  main_fill_kernel<<<grid, 256, 0, g->stream>>>(d_pay2 + e_cov, e_main, g->d_src_gather, d_main_src);
  LUXB_TRY(tmp.alloc(&d_main_row, (uint64_t)g->n_part + 4));
  main_indeg_kernel<<<grid, 256, 0, g->stream>>>(g->d_row_end, g->n_part, d_flag, d_hub_idx, d_cov, d_main_row);
  LUXB_CUDA(cudaGetLastError());
  tb = 0;
  LUXB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, d_main_row, d_main_row, (int)g->n_part, g->stream));
  LUXB_TRY(tmp.alloc((char**)&d_scan_tmp, tb + 256));
  LUXB_CUDA(cub::DeviceScan::InclusiveSum(d_scan_tmp, tb, d_main_row, d_main_row, (int)g->n_part, g->stream));
  uint64_t chk = 0;
  LUXB_CUDA(cudaMemcpyAsync(&chk, d_main_row + g->n_part - 1, 8, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  if (chk != e_main) { set_error("panel split: offsets do not add up"); return LUXB_ERR_STATE; }
  tmp.release(d_scan_tmp);
  tmp.release(d_key2);
  tmp.release(d_pay2);
  const SegShapeInfo mshp = kSegMainInfo[g->seg_main_shape];
  StreamBlocks mblk{};
  mblk.n_blocks = 1;
  mblk.vfirst[0] = 0; mblk.vfirst[1] = g->n_part;
  mblk.ebase[0] = 0; mblk.ebase[1] = e_main;
  LUXB_TRY((build_seg_stream<uint32_t, uint32_t>(g, g->sb_main, d_main_row, g->n_part, d_main_src, e_main, mblk, (uint32_t)mshp.stage_edges,
                                                 (uint32_t)mshp.piece, 0, true, d_hub_bits, nullptr)));
  tmp.release(d_main_row);
  tmp.release(d_main_src);

  // 6. raw panel sums: one slot per (block, hub); slots of (block, hub) pairs without edges stay 0 forever
  // (the program's identity: 0 bits for sums and max labels, all ones for min distances)
  LUXB_TRY(dmalloc(&g->d_sb_partial, (uint64_t)NV + 1));
  LUXB_CUDA(cudaMemsetAsync(g->d_sb_partial, g->cfg.app == LUXB_SSSP ? 0xFF : 0, ((size_t)NV + 1) * 4, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  g->d_hub_vtx = d_hub_vtx; tmp.keep(d_hub_vtx);
  g->d_hub_bits = d_hub_bits; tmp.keep(d_hub_bits);
  g->sb_n_hub = Nh;
  g->sb_n_blocks = NB;
  g->sb_bs = bs;
  g->sb_n_src = n_src;
  g->sb_on = true;
  g->stats.panel_edges = e_cov;
  g->stats.panel_hubs = Nh;
  g->stats.panel_blocks = NB;
  if (g->cfg.verbose)
    printf("[luxb rank %d] source-blocked sweep: %u hub destinations (in-degree >= %u) x %u blocks of %u hot sources; panel %llu edges "
           "(%.1f %%), main %llu edges\n", g->cfg.rank, Nh, min_indeg, NB, bs, (unsigned long long)e_cov, 100.0 * e_cov / g->e_part,
           (unsigned long long)e_main);
  return 0;
}

// The pull sweep's structures (PageRank; CC / SSSP pull iterations): the flagged stream(s) of seg.cuh.
// LUXB_SWEEP=merge keeps the merge-path tiles of pull.cuh.
static int build_seg_sweep(luxb_graph* g) {
  g->seg_on = false;
  g->sb_on = false;
  if (const char* env = getenv("LUXB_SWEEP")) if (!strcmp(env, "merge")) return 0;
  if (g->n_part == 0 || g->cfg.zero_copy_edges || g->e_part >= 0xFFFFFFF0ull) return 0;  // zero-copy graphs keep the canonical arrays
  g->seg_main_shape = env_int("LUXB_SEG_MAIN_SHAPE", 6);
  if (g->seg_main_shape < 0 || g->seg_main_shape >= kNumSegMain) g->seg_main_shape = 0;
  g->seg_panel_shape = env_int("LUXB_SEG_PANEL_SHAPE", 1);
  if (g->seg_panel_shape < 0 || g->seg_panel_shape >= kNumSegPanel) g->seg_panel_shape = 0;
  LUXB_TRY(build_panel_layout(g));
  if (!g->sb_on) {
    free_layout(g->sb_panel);
    LUXB_TRY(build_plain_seg_layout(g));
  }
  g->seg_on = true;
  return 0;
}

extern "C++" {
template <class Prog, class Shape>
static int launch_seg_shape(luxb_graph* g, const SegArgs<Prog>& a, int ctas_per_sm) {
  auto kern = seg_tile_kernel<Prog, Shape>;
  // The panel kernel holds one persistent CTA with ~all of the shared memory on every SM it runs on; when the cold half
  // of the exchange is in flight on the second stream (several ranks), a few SMs are left to NCCL's channel CTAs —
  // otherwise the collective could only start once the panel kernel has finished.
  const int reserve = (Shape::kPanel && g->packed) ? g->panel_reserve_sms : 0;
  LUXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape::kSmemBytes));
  int carve_pct = (int)std::min<size_t>(100, (ctas_per_sm * (Shape::kSmemBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
  LUXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve_pct));
  const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)std::max(g->num_sms - reserve, 1) * ctas_per_sm, a.n_stages);
  kern<<<grid, Shape::kThreads, Shape::kSmemBytes, g->stream>>>(a);
  LUXB_CUDA(cudaGetLastError());
  return 0;
}

template <class Prog>
static void fill_seg_args(SegArgs<Prog>& a, const PullLayout& L) {
  fill_fixup_args(a.p, L);
  a.p.close_vtx = L.d_close;
  a.words = L.d_src;
  a.n_stages = L.n_stages;
}
}  // extern "C++"

extern "C++" {
// main (L1) stream of a pull sweep: seg kernel + fix-up + vertices without in-edges.
// out_buffer >= 0 (PageRank): index of the value buffer written, for the once-per-buffer constants of edge-less vertices;
// < 0 (labels): `out_local` already holds the old values, edge-less vertices keep them.
template <class Prog>
static int launch_seg_main(luxb_graph* g, PullLayout& L, const typename Prog::Vertex* x_nat, const typename Prog::Vertex* x_cold,
                           typename Prog::Vertex* out_local, int out_buffer, const typename Prog::Params& prm, const uint32_t* hub_bits) {
  SegArgs<Prog> a{};
  fill_seg_args(a, L);
  a.p.n_part = L.n_vtx;
  a.p.row_left = g->row_left;
  a.p.x_nat = x_nat;
  a.p.x_old = x_cold;
  a.p.x_hot = reinterpret_cast<const typename Prog::Vertex*>(g->d_hot);
  a.p.hot_n = g->hot_n;
  a.p.out = out_local;
  a.p.prm = prm;
  a.p.hub_bits = hub_bits;
  a.p.l2_hints = g->l2_hints;
  LUXB_TRY(wait_cold_exchange(g));
  a.tile_counter = reinterpret_cast<uint32_t*>(g->d_counters + 2);
  LUXB_CUDA(cudaMemsetAsync(a.tile_counter, 0, 4, g->stream));
  switch (g->seg_main_shape) {
#define LUXB_CASE_MSHAPE(id, warps, stages, rounds) \
    case id: LUXB_TRY((launch_seg_shape<Prog, SegMain##id>(g, a, g->pull_ctas))); break;
    LUXB_SEG_MAIN_SHAPES(LUXB_CASE_MSHAPE)
    default: set_error("bad seg shape"); return LUXB_ERR_STATE;
  }
  g->stats.kernel_launches++;
  pt_mark(g, 0);
  LUXB_TRY(launch_fixup(g, a.p, L));
  // vertices without in-edges in this stream.  PageRank: update(identity) is a constant -> written once per value
  // buffer.  Hubs among them need the raw identity every sweep (the combine overwrites it).
  if (out_buffer >= 0 && L.n_empty && !g->empties_done[out_buffer]) {
    empties_kernel<Prog><<<grid_for(L.n_empty, 256, g->num_sms * 8), 256, 0, g->stream>>>(a.p, L.d_empty, L.n_empty);
    g->empties_done[out_buffer] = true;
    g->stats.kernel_launches++;
  }
  if (L.n_empty_hub) {
    empties_kernel<Prog><<<grid_for(L.n_empty_hub, 256, g->num_sms * 8), 256, 0, g->stream>>>(a.p, L.d_empty_hub, L.n_empty_hub);
    g->stats.kernel_launches++;
  }
  LUXB_CUDA(cudaGetLastError());
  pt_mark(g, 1);
  return 0;
}

// one pull sweep = [panel stream (shared-memory gathers) +] main stream (L1 gathers) [+ hub combine]
template <class Prog>
static int sweep_seg(luxb_graph* g, const typename Prog::Vertex* x_nat, const typename Prog::Vertex* x_cold, typename Prog::Vertex* out_local,
                     int out_buffer, const typename Prog::Params& prm) {
  using Acc = typename Prog::Acc;
  LUXB_TRY(kt_begin(g));
  if (g->sb_on) {
    PullLayout& PL = g->sb_panel;
    SegArgs<Prog> pa{};
    fill_seg_args(pa, PL);
    pa.p.x_hot = reinterpret_cast<const typename Prog::Vertex*>(g->d_hot);
    pa.p.out = reinterpret_cast<typename Prog::Vertex*>(g->d_sb_partial);
    pa.p.raw_out = 1;
    pa.p.prm = prm;
    pa.bs = g->sb_bs;
    pa.n_blocks = g->sb_n_blocks;
    for (uint32_t b = 0; b < g->sb_n_blocks; ++b) pa.super_end[b] = g->sb_super_end[b];
    pa.tile_counter = reinterpret_cast<uint32_t*>(g->d_counters + 4);
    LUXB_CUDA(cudaMemsetAsync(pa.tile_counter, 0, 4, g->stream));
    switch (g->seg_panel_shape) {
#define LUXB_CASE_PSHAPE(id, warps, stages, rounds, tab, v) \
      case id: LUXB_TRY((launch_seg_shape<Prog, SegPanel##id>(g, pa, 1))); break;
      LUXB_SEG_PANEL_SHAPES(LUXB_CASE_PSHAPE)
      default: set_error("bad panel shape"); return LUXB_ERR_STATE;
    }
    g->stats.kernel_launches++;
    pt_mark(g, 5);
    LUXB_TRY(launch_fixup(g, pa.p, PL));
    pt_mark(g, 1);
  }
  LUXB_TRY((launch_seg_main<Prog>(g, g->sb_main, x_nat, x_cold, out_local, out_buffer, prm, g->sb_on ? g->d_hub_bits : nullptr)));
  LUXB_TRY(kt_end(g));
  if (g->sb_on) {
    CombineArgs<Prog> ca{};
    ca.hub_vtx = g->d_hub_vtx;
    ca.n_hub = g->sb_n_hub;
    ca.n_blocks = g->sb_n_blocks;
    ca.row_left = g->row_left;
    ca.pb = g->sb_pb;
    ca.partial = reinterpret_cast<const Acc*>(g->d_sb_partial);
    ca.x_nat = x_nat;
    ca.out = out_local;
    ca.prm = prm;
    combine_hub_kernel<Prog><<<grid_for(g->sb_n_hub, 256, g->num_sms * 8), 256, 0, g->stream>>>(ca);
    LUXB_CUDA(cudaGetLastError());
    g->stats.kernel_launches++;
    pt_mark(g, 6);
  }
  return 0;
}
}  // extern "C++"

// After a sweep (or luxb_set_values): x_new holds this rank's final slice in natural order.  Make it visible to the
// next sweep of every rank: refresh the hot copies and, on several ranks, run the PACKED exchange — only vertices that are
// ever gathered travel (build.cuh), as one balanced all-gather written as kernels over peer memory: pack_push_kernel stores
// every owned entry into the transfer array of the rank holding its EQUAL chunk, a barrier (flag kernel) says the pushes
// have landed, chunk_pull_kernel copies the other ranks' chunks over NVLink.  Equal chunks because edge- or cost-balanced
// partitions own very different numbers of vertices (RMAT-27 at 8 GPUs: rank 7 owns 40 %): owner broadcasts or direct
// owner pushes (LUXB_PUSH=direct, measured 10 % slower at 4 GPUs) are bound by the biggest owner's egress.
static int pagerank_publish(luxb_graph* g, float* x_new) {
  const int me = g->cfg.rank;
  const int grid = g->num_sms * 8;
  if (g->P == 1 || !g->packed) {
    if (g->P > 1) { LUXB_TRY(allgather_slices(g, x_new, 4)); pt_mark(g, 3); }  // graphs without a hot set (tiny): plain all-gather
    if (g->hot_n) {
      hot_refresh_kernel<float><<<grid_for(g->hot_n, 256, grid), 256, 0, g->stream>>>((float*)g->d_hot, x_new, g->d_hot_order, 0, g->hot_n);
      LUXB_CUDA(cudaGetLastError());
      g->stats.kernel_launches++;
      pt_mark(g, 2);
    }
    return 0;
  }
  LUXB_TRY(wait_cold_exchange(g));  // back-to-back publishes (set_values after init): the buffer about to be packed is quiescent
  float* XTn = g->d_xt[1 - g->cur_xt];
  const uint32_t H = g->hot_n;
  const uint64_t Ch = g->xt_hot_chunk, Cc = g->xt_cold_chunk, cold_base = Ch * g->P;  // XT = [P hot chunks | P cold chunks]
  const uint32_t nh_me = g->hot_off[me + 1] - g->hot_off[me], nc_me = g->cold_off[me + 1] - g->cold_off[me];
  const bool p2p = g->p2p_ready && g->cfg.exchange != LUXB_EXCHANGE_NCCL;
  if (p2p) {
    // ---- balanced all-gather with three kernel launches and one 4-byte collective per iteration ----
    // 1. pack + push: every owned entry goes straight to the rank holding its equal chunk (remote stores);
    // 2. barrier: all pushes have landed;
    // 3. pull: every rank copies the chunks it does not hold from their holders (peer loads) — the hot region on the
    //    compute stream (the next sweep's panel gather needs it first), the cold region on the second stream,
    //    overlapped with that panel kernel (only the next MAIN sweep waits for it).
    if (nh_me + nc_me) {
      PackPushArgs<float> pa{};
      pa.x_local = x_new + g->row_left;
      pa.list = g->d_pack_list;
      pa.n_hot = nh_me;
      pa.n_cold = nc_me;
      pa.hot_pos0 = g->hot_off[me];
      pa.cold_pos0 = g->cold_off[me];
      pa.hot_chunk = Ch;
      pa.cold_chunk = Cc;
      pa.cold_base = cold_base;
      for (int k = 0; k < g->P; ++k) pa.xt[k] = reinterpret_cast<float*>(g->peer_xt[1 - g->cur_xt][k]);
      pa.P = g->P;
      const int pgrid = grid_for((uint64_t)nh_me + nc_me, 256, grid);
      if (g->direct_push) pack_push_kernel<float, true><<<pgrid, 256, 0, g->stream>>>(pa);
      else pack_push_kernel<float, false><<<pgrid, 256, 0, g->stream>>>(pa);
      LUXB_CUDA(cudaGetLastError());
      g->stats.kernel_launches++;
    }
    pt_mark(g, 7);
    LUXB_TRY(p2p_barrier(g));
    pt_mark(g, 4);
    auto pull = [&](uint64_t base, uint64_t C, cudaStream_t st, int ctas) -> int {
      ChunkPullArgs ca{};
      ca.dst = XTn + base;
      for (int k = 0; k < g->P; ++k) ca.src[k] = reinterpret_cast<const float*>(g->peer_xt[1 - g->cur_xt][k]) + base;
      ca.chunk = C;
      ca.P = g->P;
      ca.me = me;
      chunk_pull_kernel<<<ctas, 512, 0, st>>>(ca);
      LUXB_CUDA(cudaGetLastError());
      g->stats.kernel_launches++;
      return 0;
    };
    if (g->direct_push) {
      // everything is already in place: the owners stored into every rank's transfer array
    } else if (g->overlap_exchange) {
      LUXB_CUDA(cudaEventRecord(g->ev_pack, g->stream));  // = "barrier passed"
      LUXB_TRY(pull(0, Ch, g->stream, g->num_sms * 2));   // first in line: the next panel kernel waits for it
      LUXB_CUDA(cudaStreamWaitEvent(g->stream2, g->ev_pack, 0));
      // beside the panel kernel: only the SMs that kernel leaves free (launch_seg_shape)
      LUXB_TRY(pull(cold_base, Cc, g->stream2, std::max(2 * g->panel_reserve_sms, 8)));
      LUXB_CUDA(cudaEventRecord(g->ev_cold, g->stream2));
      g->cold_pending = true;
    } else {
      LUXB_TRY(pull(0, Ch, g->stream, g->num_sms * 2));
      LUXB_TRY(pull(cold_base, Cc, g->stream, g->num_sms * 2));
    }
    pt_mark(g, 8);
  } else {
    // NCCL only (no peer mappings): local pack, then one grouped set of in-place broadcasts per region
    if (nh_me + nc_me) {
      pack_values_kernel<float><<<grid_for((uint64_t)nh_me + nc_me, 256, grid), 256, 0, g->stream>>>(
          x_new + g->row_left, g->d_pack_list, nh_me, nc_me, XTn + g->hot_off[me], XTn + cold_base + g->cold_off[me]);
      LUXB_CUDA(cudaGetLastError());
      g->stats.kernel_launches++;
    }
    pt_mark(g, 7);
    LUXB_NCCL(nccl().GroupStart());
    for (int p = 0; p < g->P; ++p) {
      const uint32_t nh = g->hot_off[p + 1] - g->hot_off[p], nc = g->cold_off[p + 1] - g->cold_off[p];
      if (nh) LUXB_NCCL(nccl().Broadcast(XTn + g->hot_off[p], XTn + g->hot_off[p], nh, ncclFloat32, p, g->comm, g->stream));
      if (nc) LUXB_NCCL(nccl().Broadcast(XTn + cold_base + g->cold_off[p], XTn + cold_base + g->cold_off[p], nc, ncclFloat32, p, g->comm,
                                         g->stream));
    }
    LUXB_NCCL(nccl().GroupEnd());
    pt_mark(g, 8);
  }
  hot_permute_kernel<float><<<grid_for(H, 256, grid), 256, 0, g->stream>>>((float*)g->d_hot, XTn, g->d_zperm, H);
  LUXB_CUDA(cudaGetLastError());
  g->stats.kernel_launches++;
  pt_mark(g, 2);
  g->cur_xt ^= 1;
  g->replica_stale = true;
  return 0;
}

// the cold values of the previous exchange must have arrived before anything gathers them
static int wait_cold_exchange(luxb_graph* g) {
  if (g->cold_pending) {
    LUXB_CUDA(cudaStreamWaitEvent(g->stream, g->ev_cold, 0));
    g->cold_pending = false;
  }
  return 0;
}

static int pagerank_iteration(luxb_graph* g) {
  pt_mark(g, -1);
  PageRankProgram::Params prm;
  prm.init_rank = (1.0f - kAlpha) / (float)g->nv;  // pagerank_gpu.cu:144
  prm.deg = g->d_deg;
  float* x_old = (float*)g->d_val[g->cur];
  float* x_new = (float*)g->d_val[1 - g->cur];
  const float* x_cold = g->packed ? g->d_xt[g->cur_xt] + g->xt_hot_chunk * g->P : x_old;
  if (g->seg_on) {
    LUXB_TRY((sweep_seg<PageRankProgram>(g, x_old, x_cold, x_new + g->row_left, 1 - g->cur, prm)));
  } else {
    LUXB_TRY(wait_cold_exchange(g));
    LUXB_TRY(launch_pull<PageRankProgram>(g, base_layout(g, g->hot_n ? g->d_src_gather : g->d_src), x_old, x_cold, (const float*)g->d_hot,
                                          g->hot_n, x_new + g->row_left, prm));
  }
  LUXB_TRY(pagerank_publish(g, x_new));
  g->cur ^= 1;
  g->stats.edges_processed += g->e_part;
  return 0;
}

static int colfilter_iteration(luxb_graph* g) {
  float* x_old = (float*)g->d_val[g->cur];
  float* x_new = (float*)g->d_val[1 - g->cur];
  if (g->n_part) {
    CfArgs a{};
    a.row_end = g->d_row_end;
    a.src = g->d_src;
    a.weight = g->d_weight;
    a.chunk_first = g->d_chunk_first;
    a.chunk_vtx = g->d_chunk_vtx;
    a.n_part = g->n_part;
    a.n_chunks = g->n_chunks;
    a.row_left = g->row_left;
    a.x_old = x_old;
    a.partial = g->d_partial;
    a.out = x_new + (size_t)g->row_left * kCfK;
    a.n_peers = 0;
    if (g->p2p_ready && g->cfg.exchange != LUXB_EXCHANGE_NCCL)
      for (int p = 0; p < g->P; ++p)
        if (p != g->cfg.rank) a.peer_out[a.n_peers++] = (float*)g->peer_val[1 - g->cur][p] + (size_t)g->row_left * kCfK;
    uint64_t warps = g->n_chunks;
    int grid = (int)std::min<uint64_t>((warps * 32 + 255) / 256, (uint64_t)g->num_sms * 8);
    cf_chunk_kernel<<<grid, 256, 0, g->stream>>>(a);
    cf_update_kernel<<<grid_for((uint64_t)g->n_part * kCfK, 256, g->num_sms * 8), 256, 0, g->stream>>>(a);
    LUXB_CUDA(cudaGetLastError());
    g->stats.kernel_launches += 2;
  }
  if (g->P > 1) {
    if (g->p2p_ready && g->cfg.exchange != LUXB_EXCHANGE_NCCL) LUXB_TRY(p2p_barrier(g));
    else LUXB_TRY(allgather_slices(g, x_new, 4 * kCfK));
  }
  g->cur ^= 1;
  g->stats.edges_processed += g->e_part;
  return 0;
}

// one PushAppTask (push_app_task_impl, components_gpu.cu:335-522) on this rank, including the exchange
extern "C++" {
template <class Prog>
static int label_iteration(luxb_graph* g) {
  const int me = g->cfg.rank;
  uint32_t* lab = reinterpret_cast<uint32_t*>(g->d_val[0]);
  // direction + representation decisions from the gathered headers (components_gpu.cu:397-416)
  uint64_t old_size = 0;
  int dense_parts = 0, sparse_parts = 0;
  for (int p = 0; p < g->P; ++p) {
    old_size += g->h_hdr[2 * p + 1];
    if (g->h_hdr[2 * p] == LUXB_DENSE_BITMAP) dense_parts++; else sparse_parts++;
  }
  bool dense_fq = dense_parts >= sparse_parts;
  const bool pull = old_size > (uint64_t)(g->nv / 16);
  if (pull) dense_fq = true;
  const uint32_t max_nodes = g->cap[me];
  unsigned char* new_slot = g->d_fq_new;
  LUXB_CUDA(cudaMemsetAsync(new_slot, 0, 8, g->stream));  // components_gpu.cu:412

  if (pull) {
    typename Prog::Params prm{0};
    if (g->hot_n) {
      hot_refresh_kernel<uint32_t><<<grid_for(g->hot_n, 256, g->num_sms * 8), 256, 0, g->stream>>>((uint32_t*)g->d_hot, lab, g->d_hot_order, 0, g->hot_n);
      g->stats.kernel_launches++;
    }
    if (g->seg_on) {
      LUXB_TRY((sweep_seg<Prog>(g, lab, lab, g->d_cur, -1, prm)));
    } else {
      LUXB_TRY(launch_pull<Prog>(g, base_layout(g, g->hot_n ? g->d_src_gather : g->d_src), lab, lab, (const uint32_t*)g->d_hot, g->hot_n,
                                 g->d_cur, prm));
    }
    g->stats.edges_processed += g->e_part;
    g->stats.pull_iterations++;
  } else if (g->n_part && old_size) {
    PushArgs a{};
    uint64_t blocks = 0;
    for (int p = 0; p < g->P; ++p) {
      a.fr[p].slot = slot_ptr(g->d_fq_all, g, p);
      a.fr[p].row_left = g->rl[p];
      a.fr[p].n_part = g->np[p];
      a.fr[p].type = g->h_hdr[2 * p];
      a.fr[p].count = std::min(g->h_hdr[2 * p + 1], g->cap[p]);
      uint32_t entries = a.fr[p].type == LUXB_DENSE_BITMAP ? (g->h_hdr[2 * p + 1] ? g->np[p] : 0) : a.fr[p].count;
      if (a.fr[p].type == LUXB_DENSE_BITMAP && g->h_hdr[2 * p + 1] == 0) a.fr[p].n_part = 0;
      blocks += (entries + kPushThreads - 1) / kPushThreads;
    }
    a.n_parts = g->P;
    a.out_end = g->d_out_end;
    a.out_dst = g->d_out_dst;
    a.lab = lab;
    a.cur = g->d_cur;
    a.row_left = g->row_left;
    a.new_sparse = dense_fq ? 0 : 1;
    a.new_count = reinterpret_cast<uint32_t*>(new_slot) + 1;
    a.new_queue = reinterpret_cast<uint32_t*>(new_slot + 8);
    a.max_nodes = max_nodes;
    a.edges_scanned = g->d_counters;
    a.big_list = reinterpret_cast<PushArgs::BigSeg*>(g->d_big_list);
    a.big_count = reinterpret_cast<uint32_t*>(g->d_counters + 3);
    a.big_capacity = g->big_capacity;
    if (blocks) {
      LUXB_CUDA(cudaMemsetAsync(a.big_count, 0, 4, g->stream));
      push_relax_kernel<Prog><<<(unsigned)blocks, kPushThreads, 0, g->stream>>>(a);
      push_big_kernel<Prog><<<g->num_sms * 4, kPushThreads, 0, g->stream>>>(a);
      LUXB_CUDA(cudaGetLastError());
      g->stats.kernel_launches += 2;
    }
  }

  const int fgrid = grid_for(g->n_part, 256, g->num_sms * 8);
  uint32_t count = 0;
  uint64_t total = 0;
  FrontierHeader my_hdr{LUXB_SPARSE_QUEUE, 0};
  const bool dev_frontier = g->P == 1 || (g->p2p_ready && g->cfg.exchange != LUXB_EXCHANGE_NCCL);
  if (dev_frontier) {
    // ---- new frontier + exchange WITHOUT host round trips: the representation rules (components_gpu.cu:462-491) run on
    // the device (push.cuh), the published slot and labels are stored straight into every rank's slot table / label
    // replica (frontier P2P push), and the host reads the P headers once, at the end of the iteration ----
    unsigned char* slot_d = dense_fq ? new_slot : g->d_fq_tmp;   // bitmap candidate
    unsigned char* slot_s = dense_fq ? g->d_fq_tmp : new_slot;   // queue candidate (filled by the push kernels)
    LUXB_CUDA(cudaMemsetAsync(dense_fq ? slot_s : slot_d, 0, 8, g->stream));
    if (!g->d_fctl) LUXB_TRY(dmalloc(&g->d_fctl, 1));
    if (dense_fq) frontier_diff_kernel<<<fgrid, 256, 0, g->stream>>>(lab + g->row_left, g->d_cur, g->n_part, slot_d);
    frontier_fix_kernel<<<1, 32, 0, g->stream>>>(slot_d, slot_s, max_nodes, dense_fq ? 1 : 0, g->d_fctl);
    frontier_d2s_if_kernel<<<fgrid, 256, 0, g->stream>>>(&g->d_fctl->need_d2s, slot_d, g->row_left, g->n_part, slot_s, max_nodes);
    frontier_diff_if_kernel<<<fgrid, 256, 0, g->stream>>>(&g->d_fctl->need_promote, lab + g->row_left, g->d_cur, g->n_part, slot_d);
    frontier_final_kernel<<<1, 32, 0, g->stream>>>(slot_d, slot_s, dense_fq ? 1 : 0, g->d_fctl);
    frontier_pack_labels_if_kernel<<<grid_for(max_nodes, 256, g->num_sms * 4), 256, 0, g->stream>>>(&g->d_fctl->final_sparse, slot_s, max_nodes,
                                                                                                  g->row_left, g->d_cur);
    LUXB_CUDA(cudaGetLastError());
    g->stats.kernel_launches += dense_fq ? 6 : 5;
    // barrier: every rank has finished the kernels that read this iteration's slots and labels
    if (g->P > 1) LUXB_TRY(p2p_barrier(g));
    FrontierPushArgs fa{};
    fa.ctl = g->d_fctl;
    fa.slot_d = slot_d;
    fa.slot_s = slot_s;
    fa.cur = g->d_cur;
    fa.n_part = g->n_part;
    fa.cap = g->cap[me];
    fa.row_left = g->row_left;
    fa.n_dst = 0;
    for (int p = 0; p < g->P; ++p) {
      unsigned char* fq_p = (p == me) ? g->d_fq_all : reinterpret_cast<unsigned char*>(g->peer_fq[p]);
      uint32_t* lab_p = (p == me) ? lab : reinterpret_cast<uint32_t*>(g->peer_val[0][p]);
      fa.dst_slot[fa.n_dst] = fq_p + g->slot_off[me];
      fa.dst_lab[fa.n_dst] = lab_p;
      fa.n_dst++;
    }
    frontier_push_kernel<<<g->num_sms * 2, 512, 0, g->stream>>>(fa);
    LUXB_CUDA(cudaGetLastError());
    g->stats.kernel_launches++;
    if (g->P > 1) LUXB_TRY(p2p_barrier(g));  // all pushes have landed
    PartTable pt{};
    pt.P = g->P;
    if (!g->d_slot_off) {
      LUXB_TRY(dmalloc(&g->d_slot_off, LUXB_MAX_PARTS));
      LUXB_CUDA(cudaMemcpyAsync(g->d_slot_off, g->slot_off, sizeof(uint64_t) * g->P, cudaMemcpyHostToDevice, g->stream));
    }
    frontier_headers_kernel<<<1, LUXB_MAX_PARTS, 0, g->stream>>>(g->d_fq_all, pt, g->d_slot_off, g->d_hdr_all);
    LUXB_CUDA(cudaMemcpyAsync(g->h_hdr, g->d_hdr_all, (size_t)g->P * 8, cudaMemcpyDeviceToHost, g->stream));
    LUXB_CUDA(cudaStreamSynchronize(g->stream));  // the iteration's one host synchronisation
    my_hdr.type = g->h_hdr[2 * me];
    my_hdr.num_nodes = count = g->h_hdr[2 * me + 1];
    for (int p = 0; p < g->P; ++p) total += g->h_hdr[2 * p + 1];
  } else {
    // ---- new frontier of this partition (components_gpu.cu:462-491) ----
    if (dense_fq) {
      if (g->n_part) {
        frontier_diff_kernel<<<fgrid, 256, 0, g->stream>>>(lab + g->row_left, g->d_cur, g->n_part, new_slot);
        g->stats.kernel_launches++;
      }
      LUXB_CUDA(cudaMemcpyAsync(g->h_scratch, new_slot, 8, cudaMemcpyDeviceToHost, g->stream));
      LUXB_CUDA(cudaStreamSynchronize(g->stream));
      count = g->h_scratch[1];
      if (count < max_nodes) {  // demote to a queue
        LUXB_CUDA(cudaMemsetAsync(g->d_fq_tmp, 0, 8, g->stream));
        if (count) {
          frontier_d2s_kernel<<<fgrid, 256, 0, g->stream>>>(new_slot, g->row_left, g->n_part, g->d_fq_tmp, max_nodes);
          g->stats.kernel_launches++;
        }
        std::swap(g->d_fq_new, g->d_fq_tmp);
        new_slot = g->d_fq_new;
        dense_fq = false;
      }
    } else {
      LUXB_CUDA(cudaMemcpyAsync(g->h_scratch, new_slot, 8, cudaMemcpyDeviceToHost, g->stream));
      LUXB_CUDA(cudaStreamSynchronize(g->stream));
      count = g->h_scratch[1];
      if (count >= max_nodes) {  // promote: rebuild as a bitmap from the label diff (count is re-derived exactly)
        dense_fq = true;
        LUXB_CUDA(cudaMemsetAsync(new_slot, 0, 8, g->stream));
        frontier_diff_kernel<<<fgrid, 256, 0, g->stream>>>(lab + g->row_left, g->d_cur, g->n_part, new_slot);
        g->stats.kernel_launches++;
        LUXB_CUDA(cudaMemcpyAsync(g->h_scratch, new_slot, 8, cudaMemcpyDeviceToHost, g->stream));
        LUXB_CUDA(cudaStreamSynchronize(g->stream));
        count = g->h_scratch[1];
      }
    }
    my_hdr = FrontierHeader{dense_fq ? LUXB_DENSE_BITMAP : LUXB_SPARSE_QUEUE, count};
    LUXB_CUDA(cudaMemcpyAsync(new_slot, &my_hdr, 8, cudaMemcpyHostToDevice, g->stream));
    if (!dense_fq && count) {
      frontier_pack_labels_kernel<<<grid_for(count, 256, g->num_sms * 4), 256, 0, g->stream>>>(new_slot, max_nodes, g->row_left,
                                                                                              g->d_cur);
      g->stats.kernel_launches++;
    }
    LUXB_CUDA(cudaGetLastError());

    // ---- exchange: headers, then payload sized by each partition's representation ----
    if (g->P > 1) {
      LUXB_NCCL(nccl().AllGather(new_slot, g->d_hdr_all, 8, ncclUint8, g->comm, g->stream));
      LUXB_CUDA(cudaMemcpyAsync(g->h_hdr, g->d_hdr_all, (size_t)g->P * 8, cudaMemcpyDeviceToHost, g->stream));
      LUXB_CUDA(cudaStreamSynchronize(g->stream));
    } else {
      g->h_hdr[0] = my_hdr.type;
      g->h_hdr[1] = my_hdr.num_nodes;
    }
    for (int p = 0; p < g->P; ++p) total += g->h_hdr[2 * p + 1];
    if (g->P > 1) {
      LUXB_NCCL(nccl().GroupStart());
      for (int p = 0; p < g->P; ++p) {
        uint32_t type = g->h_hdr[2 * p], cnt = g->h_hdr[2 * p + 1];
        unsigned char* dst_slot = slot_ptr(g->d_fq_all, g, p);
        const unsigned char* send_slot = p == me ? new_slot : dst_slot;
        LUXB_NCCL(nccl().Broadcast(send_slot, dst_slot, 8, ncclUint8, p, g->comm, g->stream));
        if (cnt == 0 || g->np[p] == 0) continue;
        if (type == LUXB_DENSE_BITMAP) {
          size_t bm_bytes = (((size_t)g->np[p] + 31) / 32) * 4;
          LUXB_NCCL(nccl().Broadcast(send_slot + 8, dst_slot + 8, bm_bytes, ncclUint8, p, g->comm, g->stream));
          const void* send_lab = p == me ? (const void*)g->d_cur : (const void*)(lab + g->rl[p]);
          LUXB_NCCL(nccl().Broadcast(send_lab, lab + g->rl[p], (size_t)g->np[p] * 4, ncclUint8, p, g->comm, g->stream));
        } else {
          size_t qoff = 8 + (size_t)g->cap[p] * 4;
          LUXB_NCCL(nccl().Broadcast(send_slot + 8, dst_slot + 8, (size_t)cnt * 4, ncclUint8, p, g->comm, g->stream));
          LUXB_NCCL(nccl().Broadcast(send_slot + qoff, dst_slot + qoff, (size_t)cnt * 4, ncclUint8, p, g->comm, g->stream));
        }
      }
      LUXB_NCCL(nccl().GroupEnd());
    } else {
      LUXB_CUDA(cudaMemcpyAsync(g->d_fq_all, new_slot, g->slot_bytes[0], cudaMemcpyDeviceToDevice, g->stream));
      if (dense_fq && count && g->n_part)
        LUXB_CUDA(cudaMemcpyAsync(lab + g->row_left, g->d_cur, (size_t)g->n_part * 4, cudaMemcpyDeviceToDevice, g->stream));
    }
  }
  for (int p = 0; p < g->P; ++p) {
    uint32_t type = g->h_hdr[2 * p], cnt = g->h_hdr[2 * p + 1];
    if (type == LUXB_SPARSE_QUEUE && cnt) {
      frontier_apply_kernel<<<grid_for(cnt, 256, g->num_sms * 4), 256, 0, g->stream>>>(slot_ptr(g->d_fq_all, g, p), g->cap[p], cnt, lab);
      g->stats.kernel_launches++;
    }
  }
  LUXB_CUDA(cudaGetLastError());
  g->stats.last_active = total;
  g->stats.last_frontier_type = my_hdr.type;
  g->trace_active.push_back(total);
  g->trace_pull.push_back(pull ? 1 : 0);
  if (g->cfg.verbose)
    printf("rowLeft(%u) activeNodes(%u) globalActive(%llu) %s\n", g->row_left, count, (unsigned long long)total, pull ? "pull" : "push");
  return 0;
}

}  // extern "C++"

static int one_iteration(luxb_graph* g) {
  switch (g->cfg.app) {
    case LUXB_PAGERANK: return pagerank_iteration(g);
    case LUXB_COLFILTER: return colfilter_iteration(g);
    case LUXB_CC: return label_iteration<MaxLabelProgram>(g);
    case LUXB_SSSP: return label_iteration<HopDistProgram>(g);
  }
  return LUXB_ERR_ARG;
}

static int finish_timed(luxb_graph* g) {
  LUXB_TRY(wait_cold_exchange(g));  // the loop time includes the overlapped part of the last exchange
  pt_flush(g);
  if (g->pt.per_call) {  // LUXB_PHASE_TIMING=2: one line per luxb_iterate call instead of one per handle
    pt_print(g);
    for (double& v : g->pt.sum) v = 0;
    g->pt.cnt = 0;
  }
  LUXB_CUDA(cudaEventRecord(g->ev_end, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  float ms = 0.f;
  LUXB_CUDA(cudaEventElapsedTime(&ms, g->ev_begin, g->ev_end));
  g->stats.loop_seconds += ms * 1e-3;
  if (g->h_barrier_err && *g->h_barrier_err) {
    set_error("iteration barrier: rank %u did not arrive within %llu s (LUXB_BARRIER_TIMEOUT_S)", *g->h_barrier_err - 1u,
              (unsigned long long)(g->barrier_timeout_ns / 1000000000ull));
    return LUXB_ERR_STATE;
  }
  for (size_t k = 0; k + 1 < g->kt_used; k += 2) {
    float kms = 0.f;
    LUXB_CUDA(cudaEventElapsedTime(&kms, g->kt_events[k], g->kt_events[k + 1]));
    g->stats.dominant_kernel_seconds += kms * 1e-3;
    g->stats.dominant_kernel_launches++;
  }
  g->kt_used = 0;
  if (g->d_out_end) {
    unsigned long long scanned = 0;
    LUXB_CUDA(cudaMemcpy(&scanned, g->d_counters, 8, cudaMemcpyDeviceToHost));
    LUXB_CUDA(cudaMemset(g->d_counters, 0, 8));
    g->stats.edges_processed += scanned;
  }
  return 0;
}

int luxb_iterate(luxb_graph* g, int iters, uint64_t* active_out) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  if (!g->inited) { set_error("luxb_iterate before luxb_init"); return LUXB_ERR_STATE; }
  LUXB_ARG(iters >= 0, "negative iteration count");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  if (const char* env = getenv("LUXB_PHASE_TIMING")) { g->pt.on = atoi(env) != 0; g->pt.per_call = atoi(env) == 2; }  // may change between calls
  LUXB_CUDA(cudaEventRecord(g->ev_begin, g->stream));
  for (int i = 0; i < iters; ++i) {
    LUXB_TRY(one_iteration(g));
    g->stats.iterations++;
  }
  LUXB_TRY(finish_timed(g));
  if (active_out) *active_out = g->stats.last_active;
  return 0;
}

int luxb_run_to_convergence(luxb_graph* g, int max_iters, int* iters_out) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  if (!g->inited) { set_error("luxb_run_to_convergence before luxb_init"); return LUXB_ERR_STATE; }
  LUXB_ARG(g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP, "only push apps converge (pagerank/col_filter run -ni iterations)");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  LUXB_CUDA(cudaEventRecord(g->ev_begin, g->stream));
  int it = 0;
  while (max_iters <= 0 || it < max_iters) {
    LUXB_TRY(one_iteration(g));
    g->stats.iterations++;
    ++it;
    if (g->stats.last_active == 0) break;  // components.cc:116-123 without the 4-deep window
  }
  LUXB_TRY(finish_timed(g));
  if (iters_out) *iters_out = it;
  return 0;
}

// PageRank on several ranks exchanges only the packed transfer array every iteration: the natural-order replica is
// completed on demand (collective: every rank must make the same call)
static int refresh_replica(luxb_graph* g) {
  if (g->cfg.app == LUXB_PAGERANK && g->P > 1 && g->replica_stale) {
    LUXB_TRY(allgather_slices(g, g->d_val[g->cur], 4));
    g->replica_stale = false;
  }
  return 0;
}

int luxb_get_values(luxb_graph* g, void* host_out, size_t bytes) {
  LUXB_ARG(g && host_out, "NULL argument");
  if (!g->inited) { set_error("luxb_get_values before luxb_init"); return LUXB_ERR_STATE; }
  size_t need = (size_t)g->nv * g->vbytes;
  LUXB_ARG(bytes == need, "buffer is %zu bytes, vertex values need %zu", bytes, need);
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  LUXB_TRY(refresh_replica(g));
  const char* srcp = (const char*)((g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP) ? g->d_val[0] : g->d_val[g->cur]);
  LUXB_CUDA(cudaMemcpyAsync(host_out, srcp, need, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  return 0;
}

int luxb_get_local_values(luxb_graph* g, void* host_out, size_t bytes) {
  LUXB_ARG(g && (host_out || g->n_part == 0), "NULL argument");
  if (!g->inited) { set_error("luxb_get_local_values before luxb_init"); return LUXB_ERR_STATE; }
  size_t need = (size_t)g->n_part * g->vbytes;
  LUXB_ARG(bytes == need, "buffer is %zu bytes, this rank's %u vertex values need %zu", bytes, g->n_part, need);
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  const char* base = (const char*)((g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP) ? g->d_val[0] : g->d_val[g->cur]);
  if (need) LUXB_CUDA(cudaMemcpyAsync(host_out, base + (size_t)g->row_left * g->vbytes, need, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  return 0;
}

// values of the vertices are in place in the current replica (whole array, or only this rank's slice): make them the
// state the next iteration starts from on every rank
static int values_installed(luxb_graph* g, bool whole_array) {
  const bool labels = g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP;
  if (g->cfg.app == LUXB_PAGERANK) {
    g->empties_done[g->cur] = false;  // caller data now sits where the constants of the edge-less vertices were
    LUXB_TRY(pagerank_publish(g, (float*)g->d_val[g->cur]));
    g->replica_stale = !whole_array && g->P > 1;
  } else {
    if (!whole_array && g->P > 1) LUXB_TRY(allgather_slices(g, labels ? g->d_val[0] : g->d_val[g->cur], g->vbytes));
    if (labels) LUXB_TRY(reset_label_state(g, true));
  }
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  return 0;
}

int luxb_set_values(luxb_graph* g, const void* host_in, size_t bytes) {
  LUXB_ARG(g && host_in, "NULL argument");
  if (!g->inited) { set_error("luxb_set_values before luxb_init"); return LUXB_ERR_STATE; }
  size_t need = (size_t)g->nv * g->vbytes;
  LUXB_ARG(bytes == need, "buffer is %zu bytes, vertex values need %zu", bytes, need);
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  const bool labels = g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP;
  char* dstp = (char*)(labels ? g->d_val[0] : g->d_val[g->cur]);
  LUXB_CUDA(cudaMemcpyAsync(dstp, host_in, need, cudaMemcpyHostToDevice, g->stream));
  return values_installed(g, true);
}

int luxb_set_local_values(luxb_graph* g, const void* host_in, size_t bytes) {
  LUXB_ARG(g && (host_in || g->n_part == 0), "NULL argument");
  if (!g->inited) { set_error("luxb_set_local_values before luxb_init"); return LUXB_ERR_STATE; }
  size_t need = (size_t)g->n_part * g->vbytes;
  LUXB_ARG(bytes == need, "buffer is %zu bytes, this rank's %u vertex values need %zu", bytes, g->n_part, need);
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  const bool labels = g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP;
  char* dstp = (char*)(labels ? g->d_val[0] : g->d_val[g->cur]);
  if (need) LUXB_CUDA(cudaMemcpyAsync(dstp + (size_t)g->row_left * g->vbytes, host_in, need, cudaMemcpyHostToDevice, g->stream));
  return values_installed(g, false);
}

int luxb_check(luxb_graph* g, uint64_t* mistakes_out) {
  LUXB_ARG(g && mistakes_out, "NULL argument");
  if (!g->inited) { set_error("luxb_check before luxb_init"); return LUXB_ERR_STATE; }
  LUXB_ARG(g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP,
           "the reference has no check for pagerank / col_filter (CHECK_TASK_ID is not registered in pull_model.inl:482-521)");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  LUXB_CUDA(cudaMemsetAsync(g->d_counters + 1, 0, 8, g->stream));
  const uint32_t* lab = reinterpret_cast<const uint32_t*>(g->d_val[0]);
  int grid = grid_for(g->n_part, 256, g->num_sms * 8);
  if (g->n_part) {
    if (g->cfg.app == LUXB_CC)
      check_kernel<MaxLabelProgram><<<grid, 256, 0, g->stream>>>(g->d_row_end, g->d_src, g->n_part, g->row_left, g->nv, lab, g->d_counters + 1);
    else
      check_kernel<HopDistProgram><<<grid, 256, 0, g->stream>>>(g->d_row_end, g->d_src, g->n_part, g->row_left, g->nv, lab, g->d_counters + 1);
    LUXB_CUDA(cudaGetLastError());
  }
  unsigned long long bad = 0;
  LUXB_CUDA(cudaMemcpyAsync(&bad, g->d_counters + 1, 8, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  *mistakes_out = bad;
  return 0;
}

int luxb_stats(const luxb_graph* g, luxb_stats_t* out) {
  LUXB_ARG(g && out, "NULL argument");
  *out = g->stats;
  return 0;
}

int luxb_trace(const luxb_graph* g, uint64_t* active, int32_t* pull, int max_entries) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  int n = (int)std::min<size_t>(g->trace_active.size(), (size_t)std::max(max_entries, 0));
  for (int i = 0; i < n; ++i) {
    if (active) active[i] = g->trace_active[i];
    if (pull) pull[i] = g->trace_pull[i];
  }
  return n;
}

int luxb_enable_kernel_timing(luxb_graph* g, int on) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  g->kernel_timing = on != 0;
  return 0;
}

int luxb_get_out_degree(luxb_graph* g, luxb_vid* host_out, size_t bytes) {
  LUXB_ARG(g && host_out, "NULL argument");
  if (!g->inited || !g->d_deg) { set_error("out-degrees exist only for an initialised PageRank graph"); return LUXB_ERR_STATE; }
  LUXB_ARG(bytes == (size_t)g->nv * 4, "buffer must hold nv u32");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  LUXB_CUDA(cudaMemcpyAsync(host_out, g->d_deg, bytes, cudaMemcpyDeviceToHost, g->stream));
  LUXB_CUDA(cudaStreamSynchronize(g->stream));
  return 0;
}

// dev tooling: raw gather rate over this partition's (possibly hot-packed) source ids, no reduction structure
__global__ void debug_gather_kernel(const uint32_t* __restrict__ idx, const float* __restrict__ nat, const float* __restrict__ hot,
                                    uint32_t hot_n, uint64_t m, float* out) {
  float acc = 0.f;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < m; base += stride * 8) {
    uint32_t id[8];
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { uint64_t i = base + k * stride; id[k] = i < m ? __ldg(idx + i) : hot_n; }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(id[k] < hot_n ? hot + id[k] : nat + (id[k] - hot_n));
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
  }
  if (acc == 123.456f) out[0] = acc;
}

int luxb_debug_gather_ms(luxb_graph* g, int packed, float* ms_out) {
  LUXB_ARG(g && ms_out && g->inited && g->cfg.app == LUXB_PAGERANK && g->P == 1, "needs an initialised single-rank PageRank graph");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  const bool use_hot = packed && g->hot_n;
  const uint32_t* idx = use_hot ? g->d_src_gather : g->d_src;
  const float* nat = (const float*)g->d_val[g->cur];
  const uint32_t hn = use_hot ? g->hot_n : 0;
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    LUXB_CUDA(cudaEventRecord(g->ev_begin, g->stream));
    debug_gather_kernel<<<g->num_sms * 4, 256, 0, g->stream>>>(idx, nat, (const float*)g->d_hot, hn, g->e_part, (float*)g->d_head);
    LUXB_CUDA(cudaEventRecord(g->ev_end, g->stream));
    LUXB_CUDA(cudaStreamSynchronize(g->stream));
    float ms = 0.f;
    LUXB_CUDA(cudaEventElapsedTime(&ms, g->ev_begin, g->ev_end));
    if (r > 0 && ms < best) best = ms;
  }
  *ms_out = best;
  return 0;
}

int luxb_device_view_get(luxb_graph* g, luxb_device_view* out) {
  LUXB_ARG(g && out, "NULL argument");
  const bool labels = g->cfg.app == LUXB_CC || g->cfg.app == LUXB_SSSP;
  out->values = g->inited ? (labels ? g->d_val[0] : g->d_val[g->cur]) : nullptr;
  out->row_end = g->d_row_end;
  out->src = g->d_src;
  out->stream = g->stream;
  out->row_left = g->row_left;
  out->row_right = g->row_left + g->n_part - 1;
  out->local_edges = g->e_part;
  return 0;
}

int luxb_get_local_csc(luxb_graph* g, luxb_eid* row_end_abs, luxb_vid* src, int32_t* weight) {
  LUXB_ARG(g != nullptr, "graph is NULL");
  LUXB_CUDA(cudaSetDevice(g->cfg.device));
  if (row_end_abs && g->n_part) {
    LUXB_CUDA(cudaMemcpy(row_end_abs, g->d_row_end, (size_t)g->n_part * 8, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < g->n_part; ++i) row_end_abs[i] += g->col_left;
  }
  if (src && g->e_part) LUXB_CUDA(cudaMemcpy(src, g->d_src, g->e_part * 4, cudaMemcpyDeviceToHost));
  if (weight) {
    LUXB_ARG(g->d_weight != nullptr, "graph has no weights");
    if (g->e_part) LUXB_CUDA(cudaMemcpy(weight, g->d_weight, g->e_part * 4, cudaMemcpyDeviceToHost));
  }
  return 0;
}

void luxb_close(luxb_graph* g) {
  if (!g) return;
  pt_print(g);
  cudaSetDevice(g->cfg.device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->stream2) cudaStreamSynchronize(g->stream2);
  if (g->d_hot) cudaCtxResetPersistingL2Cache();  // release the lines pinned for the hot copies
  p2p_unmap(g);
  if (g->stream2) cudaStreamSynchronize(g->stream2);
  if (g->comm) nccl().CommDestroy(g->comm);
  if (g->ev_pack) cudaEventDestroy(g->ev_pack);
  if (g->ev_cold) cudaEventDestroy(g->ev_cold);
  if (g->stream2) cudaStreamDestroy(g->stream2);

  void* ptrs[] = {g->d_row_end, g->d_row_end32, g->d_src, g->d_weight, g->d_tile_v, g->d_head, g->d_tail, g->d_deg, g->d_val[0], g->d_val[1],
                  g->d_cur, g->d_out_end, g->d_out_dst, g->d_fq_all, g->d_fq_new, g->d_fq_tmp, g->d_hdr_all, g->d_counters,
                  g->d_chunk_first, g->d_chunk_vtx, g->d_partial, g->d_sync, g->d_hot_order, g->d_src_gather,
                  g->d_carry, g->d_carry_flag, g->d_block_agg, g->d_block_flag, g->d_hot, g->d_big_list};
  for (void* p : ptrs) {
    if (!p) continue;
    if (std::find(g->host_allocs.begin(), g->host_allocs.end(), p) != g->host_allocs.end()) cudaFreeHost(p);
    else cudaFree(p);
  }
  free_layout(g->sb_main);
  free_layout(g->sb_panel);
  if (g->base_view.d_chain) cudaFree(g->base_view.d_chain);
  if (g->d_hub_vtx) cudaFree(g->d_hub_vtx);
  if (g->d_hub_bits) cudaFree(g->d_hub_bits);
  if (g->d_sb_partial) cudaFree(g->d_sb_partial);
  for (void* q : {(void*)g->d_zperm, (void*)g->d_pack_list, (void*)g->d_xt[0], (void*)g->d_xt[1], (void*)g->d_fctl, (void*)g->d_slot_off, (void*)g->d_flags})
    if (q) cudaFree(q);
  if (g->h_barrier_err) cudaFreeHost(g->h_barrier_err);
  if (g->h_hdr) cudaFreeHost(g->h_hdr);
  if (g->h_scratch) cudaFreeHost(g->h_scratch);
  for (cudaEvent_t e : g->kt_events) cudaEventDestroy(e);
  if (g->ev_begin) cudaEventDestroy(g->ev_begin);
  if (g->ev_end) cudaEventDestroy(g->ev_end);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

}  // extern "C"
