// pull.cuh — pull-model gather over a partition's CSC slice (replaces pr_kernel pagerank_gpu.cu:49-102 and
// cc_pull_kernel / sssp_pull_kernel components_gpu.cu:85-130, sssp_gpu.cu:85-130).
//
// Design (B200-first, not a translation): the partition's work list is the MERGE of its nPart vertex-end markers
// (row_end) and its ePart in-edges; it is cut into equal tiles of kTile merge items (merge-path), so every CTA
// gets the same amount of (vertex + edge) work no matter how skewed the in-degrees are.  Per tile:
//   1. one elected thread issues two TMA bulk copies (cp.async.bulk -> UBLKCP) that bring the tile's row_end slice
//      and source-id slice into shared memory behind an mbarrier, kStages tiles ahead, L2 evict-first (streamed);
//   2. all threads gather x_old[src] (read-only path, kIPT independent loads in flight per thread) and overwrite
//      the source ids in shared memory with the gathered contributions;
//   3. every thread walks kIPT consecutive merge items serially (conflict-free smem stride since kIPT is odd),
//      producing complete per-vertex reductions plus one leading and one trailing partial; a fixed-shape
//      segmented scan (warp shuffles + one smem hop) stitches partials across threads -> deterministic sums;
//   4. a coalesced pass applies the vertex program's update() and stores the new values — to this GPU's replica
//      and, in P2P exchange mode, straight into every peer GPU's replica (fused compute + all-gather).
// Vertices whose edge list crosses a tile boundary are finished by pull_fixup_kernel from per-tile head/tail
// partials in ascending tile order (fp64 for PageRank), so results do not depend on the grid size.
#pragma once
#include "common.cuh"
#include "programs.cuh"

namespace luxb {

template <int kThreads_, int kIPT_, int kStages_>
struct PullShape {
  static constexpr int kThreads = kThreads_;
  static constexpr int kIPT = kIPT_;
  static constexpr int kStages = kStages_;
  static constexpr int kTile = kThreads * kIPT;
  static constexpr int kAElems = kTile + 4;  // u64 row_end entries per stage (alignment slack + peek)
  static constexpr int kEElems = kTile + 8;  // u32 source ids per stage (alignment slack)
  static constexpr int kWarps = kThreads / 32;
  static constexpr size_t kStageBytes = (size_t)kAElems * 8 + (size_t)kEElems * 4;
  static constexpr size_t kSmemBytes = kStages * kStageBytes + (size_t)(kTile + 4) * 4 /*sums*/ + 64 /*scan*/ * 4 +
                                       kStages * 8 /*mbarriers*/ + 16;
  static_assert(kIPT % 2 == 1, "kIPT must be odd: thread-contiguous smem walks are then bank-conflict free");
  static_assert(kThreads % 32 == 0 && kWarps <= 16, "");
};

template <class Prog>
struct PullArgs {
  const uint64_t* row_end;   // [nPart + 4] end offsets relative to the partition's first edge; padded with ~0
  const uint32_t* src;       // [ePart + 8] source vertex ids (global)
  const uint32_t* tile_v;    // [nTiles + 1] merge-path split: vertices consumed before each tile
  uint32_t n_part;           // vertices in this partition
  uint64_t e_part;           // edges in this partition
  uint32_t n_tiles;
  uint32_t row_left;         // global id of local vertex 0
  const typename Prog::Vertex* x_old;  // [nv] replica of last iteration's values (global index)
  typename Prog::Vertex* out;          // [nPart] this partition's new values (local index)
  typename Prog::Acc* head_partial;    // [nTiles]
  typename Prog::Acc* tail_partial;    // [nTiles]
  typename Prog::Params prm;
  int n_peers;                                        // P2P exchange: peers' slice pointers (local index)
  typename Prog::Vertex* peer_out[LUXB_MAX_PEERS];
};

__global__ void tile_table_kernel(const uint64_t* __restrict__ row_end, uint32_t n_part, uint64_t e_part, uint32_t tile,
                                  uint32_t n_tiles, uint32_t* __restrict__ tile_v) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  uint64_t total = (uint64_t)n_part + e_part;
  uint64_t d = (uint64_t)t * tile;
  if (d > total) d = total;
  uint64_t lo = d > e_part ? d - e_part : 0, hi = d < n_part ? d : n_part;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (row_end[mid] <= d - 1 - mid) lo = mid + 1; else hi = mid;
  }
  tile_v[t] = (uint32_t)lo;
}

template <class Prog, class Shape>
__global__ void __launch_bounds__(Shape::kThreads) pull_tile_kernel(const __grid_constant__ PullArgs<Prog> a) {
  using Acc = typename Prog::Acc;
  using Vertex = typename Prog::Vertex;
  static_assert(sizeof(Acc) == 4 && sizeof(Vertex) == 4, "4-byte vertex values");
  constexpr int kThreads = Shape::kThreads, kIPT = Shape::kIPT, kStages = Shape::kStages, kTile = Shape::kTile;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* a_buf = reinterpret_cast<uint64_t*>(smem_raw);                                         // kStages x kAElems
  uint32_t* e_buf = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kStages * Shape::kAElems * 8);  // kStages x kEElems
  Acc* sums = reinterpret_cast<Acc*>(e_buf + (size_t)kStages * Shape::kEElems);                    // kTile + 4
  Acc* scan_v = sums + (kTile + 4);                                                                // 32
  uint32_t* scan_f = reinterpret_cast<uint32_t*>(scan_v + 32);                                     // 32
  uint64_t* full = reinterpret_cast<uint64_t*>(scan_f + 32);                                       // kStages

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t total = (uint64_t)a.n_part + a.e_part;
  const uint64_t policy = l2_policy_evict_first();

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](uint32_t t, int s) {  // called by thread 0 only
    uint32_t i0 = __ldg(a.tile_v + t), i1 = __ldg(a.tile_v + t + 1);
    uint64_t d0 = (uint64_t)t * kTile, d1 = d0 + kTile < total ? d0 + kTile : total;
    uint64_t j0 = d0 - i0, j1 = d1 - i1;
    uint32_t is = i0 & ~1u;
    uint32_t bytes_a = ((i1 - is + 1) * 8 + 15) & ~15u;
    uint64_t js = j0 & ~3ull;
    uint32_t bytes_e = (uint32_t)(((j1 - js) * 4 + 15) & ~15ull);
    if (j1 == j0) bytes_e = 0;
    mbar_arrive_expect_tx(&full[s], bytes_a + bytes_e);
    bulk_g2s(a_buf + (size_t)s * Shape::kAElems, a.row_end + is, bytes_a, &full[s], policy);
    if (bytes_e) bulk_g2s(e_buf + (size_t)s * Shape::kEElems, a.src + js, bytes_e, &full[s], policy);
  };

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      uint64_t t = (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x;
      if (t < a.n_tiles) issue((uint32_t)t, s);
    }
  }

  uint32_t n = 0;
  for (uint64_t t64 = blockIdx.x; t64 < a.n_tiles; t64 += gridDim.x, ++n) {
    const uint32_t t = (uint32_t)t64;
    const int s = n % kStages;
    const uint32_t parity = (n / kStages) & 1u;
    const uint32_t i0 = __ldg(a.tile_v + t), i1 = __ldg(a.tile_v + t + 1);
    const uint64_t d0 = (uint64_t)t * kTile, d1 = d0 + kTile < total ? d0 + kTile : total;
    const uint64_t j0 = d0 - i0, j1 = d1 - i1;
    const uint32_t n_v = i1 - i0, n_e = (uint32_t)(j1 - j0), n_items = n_v + n_e;
    const uint64_t* A = a_buf + (size_t)s * Shape::kAElems + (i0 & 1u);
    uint32_t* E = e_buf + (size_t)s * Shape::kEElems + (uint32_t)(j0 & 3ull);
    Acc* vals = reinterpret_cast<Acc*>(E);

    mbar_wait(&full[s], parity);

    // ---- phase 1: gather (compute()): contributions of the tile's in-edges, kIPT loads in flight per thread ----
    {
      Acc val[kIPT];
#pragma unroll
      for (int k = 0; k < kIPT; ++k) {
        uint32_t idx = tid + k * kThreads;
        if (idx < n_e) val[k] = Prog::gather(__ldg(a.x_old + E[idx]));
      }
#pragma unroll
      for (int k = 0; k < kIPT; ++k) {
        uint32_t idx = tid + k * kThreads;
        if (idx < n_e) vals[idx] = val[k];
      }
    }
    __syncthreads();

    // ---- phase 2: serial merge walk over kIPT items per thread ----
    uint32_t d = tid * kIPT;
    if (d > n_items) d = n_items;
    uint32_t dend = d + kIPT < n_items ? d + kIPT : n_items;
    uint32_t lo = d > n_e ? d - n_e : 0, hi = d < n_v ? d : n_v;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (A[mid] <= j0 + (d - 1 - mid)) lo = mid + 1; else hi = mid;
    }
    uint32_t i = lo, j = d - lo;
    Acc acc = Prog::identity();
    bool has_c = false;
    uint32_t first_i = 0;
    Acc first_val = Prog::identity();
    uint64_t aend = A[i];
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
      if (d + k < dend) {
        if (aend <= j0 + j) {  // vertex i has no more in-edges: its reduction is complete
          if (!has_c) { has_c = true; first_i = i; first_val = acc; } else { sums[i] = acc; }
          acc = Prog::identity();
          ++i;
          aend = A[i];
        } else {
          acc = Prog::combine(acc, vals[j]);
          ++j;
        }
      }
    }
    // segmented inclusive scan of (has_c, trailing partial) over threads: warp shuffles, then one smem hop
    Acc sv = acc;
    uint32_t sf = has_c ? 1u : 0u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      Acc pv = __shfl_up_sync(0xffffffffu, sv, off);
      uint32_t pf = __shfl_up_sync(0xffffffffu, sf, off);
      if (lane >= off) {
        if (!sf) sv = Prog::combine(pv, sv);
        sf |= pf;
      }
    }
    Acc ex_v = __shfl_up_sync(0xffffffffu, sv, 1);
    uint32_t ex_f = __shfl_up_sync(0xffffffffu, sf, 1);
    if (lane == 0) { ex_v = Prog::identity(); ex_f = 0; }
    if (lane == 31) { scan_v[warp] = sv; scan_f[warp] = sf; }
    fence_proxy_async_smem();
    __syncthreads();  // all generic accesses to this stage's buffers are done -> stage can be refilled
    if (tid == 0) {
      uint64_t tn = t64 + (uint64_t)kStages * gridDim.x;
      if (tn < a.n_tiles) issue((uint32_t)tn, s);
    }
    {
      Acc pv = Prog::identity();
      uint32_t pf = 0;
      for (int w = 0; w < warp; ++w) {
        if (scan_f[w]) { pv = scan_v[w]; pf = 1; } else { pv = Prog::combine(pv, scan_v[w]); }
      }
      if (!ex_f) ex_v = Prog::combine(pv, ex_v);
      ex_f |= pf;
    }
    if (has_c) sums[first_i] = Prog::combine(ex_v, first_val);
    Acc tail = Prog::identity();
    if (tid == 0) {
      for (int w = 0; w < Shape::kWarps; ++w) {
        if (scan_f[w]) tail = scan_v[w]; else tail = Prog::combine(tail, scan_v[w]);
      }
    }
    __syncthreads();  // sums[] complete

    // ---- phase 3: update() + coalesced stores (own replica and, in P2P mode, every peer's replica) ----
    if (tid == 0) {
      a.tail_partial[t] = tail;
      if (n_v > 0) a.head_partial[t] = sums[0];
    }
    for (uint32_t li = tid; li < n_v; li += kThreads) {
      if (li == 0 && t != 0) continue;  // may continue from previous tiles: finished by pull_fixup_kernel
      uint32_t v = i0 + li;
      Vertex oldv = Prog::kNeedsOld ? __ldg(a.x_old + a.row_left + v) : Vertex();
      Vertex nv_ = Prog::update(a.row_left + v, sums[li], oldv, a.prm);
      a.out[v] = nv_;
      for (int p = 0; p < a.n_peers; ++p) a.peer_out[p][v] = nv_;
    }
    // next tile's post-gather __syncthreads orders these sums[] reads before its merge-phase writes
  }
}

// One thread per tile t >= 1 that completes at least one vertex: local vertex tile_v[t] may have started in earlier
// tiles.  Combine their tail partials in ascending tile order with this tile's head partial, then update().
template <class Prog>
__global__ void pull_fixup_kernel(const __grid_constant__ PullArgs<Prog> a) {
  using Wide = typename Prog::Wide;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (t >= a.n_tiles) return;
  uint32_t i0 = a.tile_v[t], i1 = a.tile_v[t + 1];
  if (i1 == i0) return;
  // walk back to the tile in which vertex i0's edge list starts
  uint32_t s = t - 1;
  while (s > 0 && a.tile_v[s + 1] == a.tile_v[s]) --s;
  Wide acc = Prog::widen(Prog::identity());
  for (uint32_t q = s; q < t; ++q) acc = Prog::wcombine(acc, Prog::widen(a.tail_partial[q]));
  acc = Prog::wcombine(acc, Prog::widen(a.head_partial[t]));
  uint32_t v = i0;
  typename Prog::Vertex oldv = Prog::kNeedsOld ? a.x_old[a.row_left + v] : typename Prog::Vertex();
  typename Prog::Vertex nv_ = Prog::update(a.row_left + v, Prog::narrow(acc), oldv, a.prm);
  a.out[v] = nv_;
  for (int p = 0; p < a.n_peers; ++p) a.peer_out[p][v] = nv_;
}

}  // namespace luxb
