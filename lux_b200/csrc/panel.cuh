// panel.cuh — source-blocked ("panel") half of the PageRank pull sweep: the split itself and the hub combine.
//
// Why.  Measured on B200 (profiles/r02_ubench_head.txt): divergent 4-byte gathers through L1 run at exactly one
// sector per cycle per SM (290 G/s) — the wall the L1 sweep sits on; the same gathers from a table in SHARED memory
// run at > 1000 G/s; distributed shared memory (ld.shared::cluster) is slower than L1 (33-180 G/s); and a
// shared-memory table inside the L1-gather kernel starves L1 of the lines its misses need (head of 48 K values:
// 150 G/s).  So the shared-memory gathers get a kernel launch of their own, in which NOTHING goes through L1:
//
//   * hub destinations  = local vertices with in-degree >= D (they own most edges of a skewed graph);
//   * hot source blocks = the hot-packed value space [0, Ns) cut into NB blocks of BS <= 32768 values (BS * 4 B fits
//     shared memory next to the streaming ring);
//   * every edge (hot source of block b -> hub destination h) moves from the partition's CSC into the PANEL: one
//     "virtual vertex" b * Nh + h per (block, hub), in that order, its in-edges stored as 15-BIT offsets into block b
//     plus a head flag — 2 B of edge stream instead of 4 B.  Blocks are padded to whole stages, so a stage never
//     straddles two blocks.
// The panel stream is swept by seg_tile_kernel<kPanel> (seg.cuh): its producer warp keeps block b's values resident in
// shared memory (one TMA bulk load of BS * 4 B from the hot copies per block change) and its gathers are ld.shared.  It
// writes RAW partial sums, one per virtual vertex.  The remaining edges stay in the "main" stream swept through L1;
// for hub vertices that sweep stores its raw sum too, and combine_hub_kernel adds main + sum_b panel partials in fp64
// (fixed order: deterministic) and applies the vertex program's update().  Replaces pr_kernel (pagerank_gpu.cu:49-102).
#pragma once
#include "common.cuh"
#include "programs.cuh"
#include "pull.cuh"

namespace luxb {

constexpr int kPanelMaxBlocks = 64;

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

// ---- per-iteration: hubs = main raw sum + panel partials, fp64, fixed order; then update() -------------------------
template <class Prog>
struct CombineArgs {
  const uint32_t* hub_vtx;   // [n_hub] local vertex ids
  uint32_t n_hub, n_blocks, row_left;
  PanelBases pb;
  const typename Prog::Acc* partial;  // [NV] raw panel reductions
  const typename Prog::Vertex* x_nat; // natural-order values of the previous iteration (update()'s old value)
  typename Prog::Vertex* out;  // [n_part] local; holds the main kernel's RAW sum for hub vertices on entry
  typename Prog::Params prm;
};

template <class Prog>
__global__ void combine_hub_kernel(const __grid_constant__ CombineArgs<Prog> a) {
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < a.n_hub; h += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = a.hub_vtx[h];
    typename Prog::Acc raw0;
    memcpy(&raw0, &a.out[v], sizeof(raw0));  // the main sweep left its RAW reduction in the value slot
    typename Prog::Wide t = Prog::widen(raw0);
    for (uint32_t b = 0; b < a.n_blocks; ++b) t = Prog::wcombine(t, Prog::widen(a.partial[a.pb.vbase[b] + h]));
    const typename Prog::Vertex oldv = Prog::kNeedsOld ? a.x_nat[a.row_left + v] : typename Prog::Vertex();
    const typename Prog::Vertex nv_ = Prog::update(a.row_left + v, Prog::narrow(t), oldv, a.prm);
    a.out[v] = nv_;
  }
}

}  // namespace luxb
