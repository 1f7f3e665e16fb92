// panel.cuh — source-blocked ("panel") half of the PageRank pull sweep.
//
// Why.  Measured on B200 (profiles/r02_ubench_head.txt): divergent 4-byte gathers through L1 run at exactly one
// sector per cycle per SM (290 G/s) — the wall pull_tile_kernel sits on; the same gathers from a table in SHARED
// memory run at > 1000 G/s; distributed shared memory (ld.shared::cluster) is slower than L1 (33-180 G/s); and a
// shared-memory table inside the L1-gather kernel starves L1 of the lines its misses need (head of 48 K values:
// 150 G/s).  So the shared-memory gathers get a kernel of their own, in which NOTHING goes through L1:
//
//   * hub destinations  = local vertices with in-degree >= D (they own most edges of a skewed graph);
//   * hot source blocks = the hot-packed value space [0, Ns) cut into NB blocks of BS values (BS * 4 B fits shared
//     memory next to the streaming ring);
//   * every edge (hot source of block b -> hub destination h) moves from the partition's CSC into the PANEL CSC:
//     one "virtual vertex" per (b, h), in (b, h) order, its in-edges stored as 16-BIT offsets into block b —
//     2 B of edge stream instead of 4 B.  Blocks are padded with edge-less virtual vertices to a multiple of the
//     super-tile, so a super-tile never straddles two blocks.
// panel_tile_kernel runs the same merge-path / warp-tile / TMA-ring machinery as pull_tile_kernel over the panel CSC;
// its producer warp additionally keeps block b's values resident in shared memory (one TMA bulk load of BS * 4 B from
// the hot copies per block change) and its gathers are ld.shared.  It writes RAW partial sums, one per virtual vertex.
// The remaining edges stay in the "main" CSC swept by pull_tile_kernel (full L1); for hub vertices it stores its raw
// sum too, and combine_hub_kernel adds main + sum_b panel partials in fp64 (fixed order: deterministic) and applies
// the vertex program's update().  Replaces, like pull.cuh, pr_kernel (pagerank_gpu.cu:49-102).
#pragma once
#include "common.cuh"
#include "programs.cuh"
#include "pull.cuh"

namespace luxb {

constexpr int kPanelMaxBlocks = 64;

template <int kIPT_, int kWarps_, int kStages_, int kTab_>
struct PanelShape {
  static constexpr int kIPT = kIPT_;
  static constexpr int kWarps = kWarps_;
  static constexpr int kThreads = 32 * (kWarps + 1);
  static constexpr int kTile = 32 * kIPT;
  static constexpr int kSuper = kTile * kWarps;
  static constexpr int kStages = kStages_;
  static constexpr int kTab = kTab_;            // capacity of the shared-memory value table (values); <= 65536
  static constexpr int kAElems = kSuper + 8;    // u32 low words of the virtual row_end
  static constexpr int kEElems = kSuper + 16;   // u16 block-local source offsets (alignment slack), multiple of 8
  static constexpr int kSumElems = kTile + 4;
  static constexpr int kHdrElems = kWarps + 6;  // super-tile id, table generation, tile_v[t0 .. t0 + kWarps]
  static constexpr size_t kSmemBytes = (size_t)kTab * 4 + (size_t)kStages * (kAElems * 4 + kEElems * 2) +
                                       (size_t)kWarps * kSumElems * 4 + (2 * kStages + 1) * 8 +
                                       (size_t)kStages * kHdrElems * 4 + 16;
  static_assert(kTab <= 65536 && kTab % 4 == 0, "block-local offsets are 16 bit");
  static_assert(kEElems % 8 == 0 && kSuper % 8 == 0, "16-byte aligned u16 stages");
};

struct PanelArgs {
  const uint32_t* row_end32;  // [NV + 8] low words of the panel CSC's end offsets
  const uint16_t* src16;      // [Ecov + 16] block-local source offsets
  const uint32_t* tile_v;     // [n_tiles + 1]
  uint32_t n_vtx;             // NV virtual vertices (all blocks, padding included)
  uint64_t e_cnt;             // Ecov
  uint32_t n_tiles;
  const float* x_hot;         // hot copies: block b's table = x_hot[b * bs, b * bs + bs)
  uint32_t bs;                // values per block (<= Shape::kTab)
  uint32_t n_blocks;
  uint32_t super_end[kPanelMaxBlocks];  // first super-tile index after block b (cumulative)
  float* out;                 // [NV] raw partial sums
  float* head_partial;        // [n_tiles]
  float* tail_partial;        // [n_tiles]
  uint32_t* tile_counter;
};

// Same tile algorithm as pull_tile_kernel (see pull.cuh for the commentary); differences are marked PANEL.
template <class Shape>
__global__ void __launch_bounds__(Shape::kThreads) panel_tile_kernel(const __grid_constant__ PanelArgs a) {
  constexpr int kIPT = Shape::kIPT, kTile = Shape::kTile, kStages = Shape::kStages, kWarps = Shape::kWarps;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tab = reinterpret_cast<float*>(smem_raw);                                   // PANEL: kTab values of block b
  uint32_t* a_buf = reinterpret_cast<uint32_t*>(tab + Shape::kTab);                  // kStages x kAElems
  uint16_t* e_buf = reinterpret_cast<uint16_t*>(a_buf + (size_t)kStages * Shape::kAElems);  // kStages x kEElems (u16)
  float* sums_all = reinterpret_cast<float*>(e_buf + (size_t)kStages * Shape::kEElems);
  uint64_t* full = reinterpret_cast<uint64_t*>(sums_all + (size_t)kWarps * Shape::kSumElems);
  uint64_t* empty = full + kStages;
  uint64_t* tab_full = empty + kStages;
  uint32_t* hdr_all = reinterpret_cast<uint32_t*>(tab_full + 1);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t total = (uint64_t)a.n_vtx + a.e_cnt;
  const uint32_t n_super = (a.n_tiles + kWarps - 1) / kWarps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
    mbar_init(tab_full, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kWarps) {
    // ===== producer warp =====
    const uint64_t policy = l2_policy_evict_first();
    uint32_t cur_b = 0, gen = 0;
    bool loaded = false;
    for (uint32_t n = 0;; ++n) {
      const int s = n % kStages;
      if (n >= (uint32_t)kStages) mbar_wait(&empty[s], ((n / kStages) - 1) & 1u);
      uint32_t T = 0;
      if (lane == 0) T = atomicAdd(a.tile_counter, 1u);
      T = __shfl_sync(0xffffffffu, T, 0);
      uint32_t* hdr = hdr_all + s * Shape::kHdrElems;
      if (T >= n_super) {
        if (lane == 0) { hdr[0] = 0xFFFFFFFFu; mbar_arrive(&full[s]); }
        break;
      }
      // PANEL: which block does this super-tile belong to (the counter is monotone, so is the block index)
      uint32_t b = cur_b;
      while (b + 1 < a.n_blocks && T >= a.super_end[b]) ++b;
      if (!loaded || b != cur_b) {
        // every consumer must be done with the super-tiles issued so far (they gather from the old table):
        // wait for the latest use of every ring stage to be released
        for (int q = 0; q < kStages; ++q) {
          if (n > (uint32_t)q) {
            const uint32_t m = n - 1 - ((n - 1 - q) % kStages);  // largest m < n with m % kStages == q
            mbar_wait(&empty[q], (m / kStages) & 1u);
          }
        }
        cur_b = b;
        loaded = true;
        ++gen;
        if (lane == 0) {
          const uint32_t bytes = a.bs * 4u;  // bs is a multiple of 4 values; the hot buffer is padded to whole blocks
          mbar_arrive_expect_tx(tab_full, bytes);
          const char* gsrc = reinterpret_cast<const char*>(a.x_hot + (size_t)b * a.bs);
          char* sdst = reinterpret_cast<char*>(tab);
          for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t chunk = bytes - off < 32768u ? bytes - off : 32768u;
            bulk_g2s(sdst + off, gsrc + off, chunk, tab_full, l2_policy_evict_last());
          }
        }
        __syncwarp();
      }
      const uint64_t t0 = (uint64_t)T * kWarps;
      if (lane <= kWarps) {
        uint64_t tt = t0 + lane < a.n_tiles ? t0 + lane : a.n_tiles;
        hdr[2 + lane] = __ldg(a.tile_v + tt);
      }
      __syncwarp();
      if (lane == 0) {
        hdr[0] = T;
        hdr[1] = gen;
        const uint64_t t1 = t0 + kWarps < a.n_tiles ? t0 + kWarps : a.n_tiles;
        const uint32_t i0 = hdr[2], i1 = hdr[2 + kWarps];
        const uint64_t d0 = t0 * kTile, d1 = t1 * kTile < total ? t1 * kTile : total;
        const uint64_t j0 = d0 - i0, j1 = d1 - i1;
        const uint32_t is = i0 & ~3u;
        const uint32_t bytes_a = ((i1 - is + 1) * 4 + 15) & ~15u;
        const uint64_t js = j0 & ~7ull;  // PANEL: 8 u16 per 16 bytes
        uint32_t bytes_e = (uint32_t)(((j1 - js) * 2 + 15) & ~15ull);
        if (j1 == j0) bytes_e = 0;
        mbar_arrive_expect_tx(&full[s], bytes_a + bytes_e);
        bulk_g2s(a_buf + (size_t)s * Shape::kAElems, a.row_end32 + is, bytes_a, &full[s], policy);
        if (bytes_e) bulk_g2s(e_buf + (size_t)s * Shape::kEElems, a.src16 + js, bytes_e, &full[s], policy);
      }
      __syncwarp();
    }
    return;
  }

  // ===== consumer warps =====
  float* sums = sums_all + (size_t)warp * Shape::kSumElems;
  uint32_t my_gen = 0;
  for (uint32_t n = 0;; ++n) {
    const int s = n % kStages;
    mbar_wait(&full[s], (n / kStages) & 1u);
    const uint32_t* hdr = hdr_all + s * Shape::kHdrElems;
    const uint32_t T = hdr[0];
    if (T == 0xFFFFFFFFu) break;
    const uint32_t gen = hdr[1];
    if (gen != my_gen) {  // PANEL: a new block's table is (being) loaded: wait for its bytes
      mbar_wait(tab_full, (gen - 1) & 1u);
      my_gen = gen;
    }
    const uint64_t t0 = (uint64_t)T * kWarps;
    const uint64_t t64 = t0 + warp;
    const bool active = t64 < a.n_tiles;
    const uint32_t t = (uint32_t)t64;
    const uint32_t si0 = hdr[2];
    const uint32_t i0 = hdr[2 + warp], i1 = hdr[3 + warp];
    const uint64_t sj0 = t0 * kTile - si0;
    const uint64_t d0 = (uint64_t)t * kTile, d1 = d0 + kTile < total ? d0 + kTile : total;
    const uint64_t j0 = d0 - i0, j1 = d1 - i1;
    const uint32_t n_v = i1 - i0, n_e = active ? (uint32_t)(j1 - j0) : 0u, n_items = n_v + n_e, j0lo = (uint32_t)j0;
    const uint32_t* A = a_buf + (size_t)s * Shape::kAElems + (si0 & 3u) + (i0 - si0);
    const uint16_t* E = e_buf + (size_t)s * Shape::kEElems + (uint32_t)(sj0 & 7ull) + (uint32_t)(j0 - sj0);

    if (active) {
      uint32_t d = lane * kIPT;
      if (d > n_items) d = n_items;
      uint32_t lo = d > n_e ? d - n_e : 0, hi = d < n_v ? d : n_v;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (A[mid] - j0lo <= d - 1 - mid) lo = mid + 1; else hi = mid;
      }
      uint32_t i = lo;
      const uint32_t j = d - lo;
      uint32_t i_next = __shfl_down_sync(0xffffffffu, i, 1);
      uint32_t j_next = __shfl_down_sync(0xffffffffu, j, 1);
      if (lane == 31) { i_next = n_v; j_next = n_e; }
      const uint32_t ne_lane = j_next - j;

      float val[kIPT];
#pragma unroll
      for (int k = 0; k < kIPT; ++k)
        if (k < (int)ne_lane) val[k] = tab[E[j + k]];  // PANEL: the gather is a shared-memory load

      float acc = 0.f;
      bool has_c = false;
      uint32_t first_i = 0;
      float first_val = 0.f;
      uint32_t aend = A[i] - j0lo;
#pragma unroll
      for (int k = 0; k < kIPT; ++k) {
        if (k < (int)ne_lane) {
          while (i < i_next && aend <= j + k) {
            if (!has_c) { has_c = true; first_i = i; first_val = acc; } else { sums[i] = acc; }
            acc = 0.f;
            ++i;
            aend = A[i] - j0lo;
          }
          acc += val[k];
        }
      }
      while (i < i_next) {
        if (!has_c) { has_c = true; first_i = i; first_val = acc; } else { sums[i] = acc; }
        acc = 0.f;
        ++i;
      }
      float sv = acc;
      uint32_t sf = has_c ? 1u : 0u;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        float pv = __shfl_up_sync(0xffffffffu, sv, off);
        uint32_t pf = __shfl_up_sync(0xffffffffu, sf, off);
        if (lane >= off) {
          if (!sf) sv = pv + sv;
          sf |= pf;
        }
      }
      float ex_v = __shfl_up_sync(0xffffffffu, sv, 1);
      if (lane == 0) ex_v = 0.f;
      if (has_c) sums[first_i] = ex_v + first_val;
      const float tail = __shfl_sync(0xffffffffu, sv, 31);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&empty[s]);
        a.tail_partial[t] = tail;
        if (n_v > 0) a.head_partial[t] = sums[0];
      }
      // PANEL: raw partial sums, no update()
      for (uint32_t li = lane; li < n_v; li += 32) {
        if (li == 0 && t != 0) continue;  // may continue from previous tiles: finished by the fix-up kernels
        a.out[i0 + li] = sums[li];
      }
      __syncwarp();
    } else {
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

// ---- one-time construction of the panel / main split --------------------------------------------------------------
__global__ void hub_flag_kernel(const uint64_t* __restrict__ row_end_rel, uint32_t n_part, uint32_t min_indeg,
                                uint32_t* __restrict__ flag) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_part; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t d = row_end_rel[i] - (i == 0 ? 0 : row_end_rel[i - 1]);
    flag[i] = d >= min_indeg ? 1u : 0u;
  }
}

// hub_idx = exclusive scan of flag.  Writes the hub list (local vertex ids, ascending) and the bitmap the main
// kernel tests (bit v of word v / 32).
__global__ void hub_list_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ hub_idx, uint32_t n_part,
                                uint32_t* __restrict__ hub_vtx, uint32_t* __restrict__ hub_bits) {
  const uint64_t n_round = ((uint64_t)n_part + 31) & ~31ull;
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_round; v += (uint64_t)gridDim.x * blockDim.x) {
    const bool f = v < n_part && flag[v];
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) hub_bits[v >> 5] = m;
    if (f) hub_vtx[hub_idx[v]] = (uint32_t)v;
  }
}

__global__ void edge_iota_kernel(uint64_t* __restrict__ payload, uint8_t* __restrict__ key, uint64_t n) {
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x) {
    payload[e] = e;
    key[e] = 255;
  }
}

// one warp per hub vertex: edges whose (hot-packed) source id lies below n_src get key = block, payload = (h, e)
__global__ void hub_key_kernel(const uint64_t* __restrict__ row_end_rel, const uint32_t* __restrict__ src_gather,
                               const uint32_t* __restrict__ hub_vtx, uint32_t n_hub, uint32_t n_src, uint32_t bs,
                               uint8_t* __restrict__ key, uint64_t* __restrict__ payload, uint32_t* __restrict__ cov_count) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t warps_total = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t h = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < n_hub; h += warps_total) {
    const uint32_t v = hub_vtx[h];
    const uint64_t b = v == 0 ? 0 : row_end_rel[v - 1], e = row_end_rel[v];
    uint32_t cnt = 0;
    for (uint64_t k = b + lane; k < e; k += 32) {
      const uint32_t id = src_gather[k];
      if (id < n_src) {
        key[k] = (uint8_t)(id / bs);
        payload[k] = (h << 32) | k;
        ++cnt;
      }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if (lane == 0) cov_count[h] = cnt;
  }
}

__global__ void key_hist_kernel(const uint8_t* __restrict__ key, uint64_t n, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_h[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  // per-CTA counts stay below 2^32: each CTA sees at most n / gridDim.x + blockDim.x keys (n < 2^32 * gridDim.x)
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(&s_h[key[e]], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (s_h[i]) atomicAdd(hist + i, (unsigned long long)s_h[i]);
}

struct PanelBases {
  uint32_t vbase[kPanelMaxBlocks + 1];  // first virtual vertex of block b; [n_blocks] = NV
};

// sorted (by block, stable) covered edges -> 16-bit offsets + per-virtual-vertex in-degree
__global__ void panel_fill_kernel(const uint8_t* __restrict__ key_sorted, const uint64_t* __restrict__ payload_sorted,
                                  uint64_t e_cov, const uint32_t* __restrict__ src_gather, uint32_t bs,
                                  const __grid_constant__ PanelBases pb, uint16_t* __restrict__ src16, uint32_t* __restrict__ vcount) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e_cov; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = key_sorted[i];
    const uint64_t p = payload_sorted[i];
    const uint32_t e = (uint32_t)p, h = (uint32_t)(p >> 32);
    src16[i] = (uint16_t)(src_gather[e] - b * bs);
    atomicAdd(vcount + pb.vbase[b] + h, 1u);
  }
}

__global__ void main_fill_kernel(const uint64_t* __restrict__ payload_sorted, uint64_t e_main, const uint32_t* __restrict__ src_gather,
                                 uint32_t* __restrict__ main_src) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e_main; i += (uint64_t)gridDim.x * blockDim.x)
    main_src[i] = src_gather[(uint32_t)payload_sorted[i]];
}

// in-degree of every local vertex inside the MAIN CSC = its in-degree minus the edges that moved to the panel
__global__ void main_indeg_kernel(const uint64_t* __restrict__ row_end_rel, uint32_t n_part, const uint32_t* __restrict__ flag,
                                  const uint32_t* __restrict__ hub_idx, const uint32_t* __restrict__ cov_count,
                                  uint64_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_part; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t d = row_end_rel[i] - (i == 0 ? 0 : row_end_rel[i - 1]);
    if (flag[i]) d -= cov_count[hub_idx[i]];
    out[i] = d;
  }
}

__global__ void pad_sentinels_kernel(uint64_t* __restrict__ row_end, uint64_t n) {
  if (blockIdx.x == 0 && threadIdx.x < 4) row_end[n + threadIdx.x] = ~0ull;
}

// ---- per-iteration: hubs = main raw sum + panel partials, fp64, fixed order; then update() -------------------------
template <class Prog>
struct CombineArgs {
  const uint32_t* hub_vtx;   // [n_hub] local vertex ids
  uint32_t n_hub, n_blocks, row_left;
  PanelBases pb;
  const float* partial;      // [NV]
  typename Prog::Vertex* out;  // [n_part] local; holds the main kernel's RAW sum for hub vertices on entry
  typename Prog::Params prm;
  int n_peers;
  typename Prog::Vertex* peer_out[LUXB_MAX_PEERS];
};

template <class Prog>
__global__ void combine_hub_kernel(const __grid_constant__ CombineArgs<Prog> a) {
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < a.n_hub; h += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = a.hub_vtx[h];
    typename Prog::Wide t = Prog::widen(a.out[v]);
    for (uint32_t b = 0; b < a.n_blocks; ++b) t = Prog::wcombine(t, Prog::widen(a.partial[a.pb.vbase[b] + h]));
    const typename Prog::Vertex nv_ = Prog::update(a.row_left + v, Prog::narrow(t), typename Prog::Vertex(), a.prm);
    a.out[v] = nv_;
    for (int p = 0; p < a.n_peers; ++p) a.peer_out[p][v] = nv_;
  }
}

}  // namespace luxb
