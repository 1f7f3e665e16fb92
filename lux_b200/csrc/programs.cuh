// programs.cuh — the vertex-program surface (init / compute / update) that Lux leaves implicit in its kernels
// (SURVEY §8b "Vertex program").  Each struct mirrors one app.h + the arithmetic of its *_gpu.cu:
//   PageRankProgram : pagerank/app.h:19-35,   pagerank_gpu.cu:86-100 (compute+update), :255-259 (init)
//   MaxLabelProgram : components/app.h:19-38, components_gpu.cu:112-122 (pull), :48-82 (push), :738-739 (init)
//   HopDistProgram  : sssp/app.h,             sssp_gpu.cu:112-122, :57-59,75-77, :733-744 (init, INF = nv)
// Kernels are templated on these; adding an app = adding a struct.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#ifndef LUXB_MAX_PEERS
#define LUXB_MAX_PEERS 63
#endif

namespace luxb {

constexpr float kAlpha = 0.15f;  // ALPHA, pagerank/app.h:24

struct PageRankProgram {
  using Vertex = float;  // stored value = rank / out-degree (pagerank_gpu.cu:98-100)
  using Acc = float;
  using Wide = double;
  struct Params {
    float init_rank;      // (1 - ALPHA) / nv, pagerank_gpu.cu:144
    const uint32_t* deg;  // global out-degrees (pull_scan_task_impl, pull_model.inl:333-343)
  };
  static constexpr bool kNeedsOld = false;
  __device__ __forceinline__ static Acc identity() { return 0.0f; }
  __device__ __forceinline__ static Acc gather(Vertex src_val) { return src_val; }
  __device__ __forceinline__ static Acc combine(Acc x, Acc y) { return x + y; }
  __device__ __forceinline__ static Wide widen(Acc x) { return (double)x; }
  __device__ __forceinline__ static Wide wcombine(Wide x, Wide y) { return x + y; }
  __device__ __forceinline__ static Acc narrow(Wide x) { return (float)x; }
  __device__ __forceinline__ static Vertex update(uint32_t v, Acc acc, Vertex, const Params& p) {
    float y = __fmaf_rn(kAlpha, acc, p.init_rank);
    uint32_t d = __ldg(p.deg + v);
    return d != 0 ? __fdiv_rn(y, (float)d) : y;
  }
};

struct MaxLabelProgram {  // connected components: label = max id that reaches the vertex
  using Vertex = uint32_t;
  using Acc = uint32_t;
  using Wide = uint32_t;
  struct Params { uint32_t unused; };
  static constexpr bool kNeedsOld = true;  // new = max(old, gathered)  (components_gpu.cu:106)
  static constexpr bool kIsMax = true;
  __device__ __forceinline__ static Acc identity() { return 0u; }
  __device__ __forceinline__ static Acc gather(Vertex src_val) { return src_val; }
  __device__ __forceinline__ static Acc combine(Acc x, Acc y) { return x > y ? x : y; }
  __device__ __forceinline__ static Wide widen(Acc x) { return x; }
  __device__ __forceinline__ static Wide wcombine(Wide x, Wide y) { return x > y ? x : y; }
  __device__ __forceinline__ static Acc narrow(Wide x) { return x; }
  __device__ __forceinline__ static Vertex update(uint32_t, Acc acc, Vertex old_v, const Params&) {
    return acc > old_v ? acc : old_v;
  }
  __device__ __forceinline__ static bool better(Vertex cand, Vertex cur) { return cand > cur; }
  __device__ __forceinline__ static Vertex atomic_relax(Vertex* addr, Vertex cand) { return atomicMax(addr, cand); }
};

struct HopDistProgram {  // the reference's "SSSP" = BFS depth (sssp_gpu.cu:122: srcLabel + 1)
  using Vertex = uint32_t;
  using Acc = uint32_t;
  using Wide = uint32_t;
  struct Params { uint32_t unused; };
  static constexpr bool kNeedsOld = true;
  static constexpr bool kIsMax = false;
  __device__ __forceinline__ static Acc identity() { return 0xFFFFFFFFu; }
  __device__ __forceinline__ static Acc gather(Vertex src_val) { return src_val + 1u; }  // INF = nv stays > nv
  __device__ __forceinline__ static Acc combine(Acc x, Acc y) { return x < y ? x : y; }
  __device__ __forceinline__ static Wide widen(Acc x) { return x; }
  __device__ __forceinline__ static Wide wcombine(Wide x, Wide y) { return x < y ? x : y; }
  __device__ __forceinline__ static Acc narrow(Wide x) { return x; }
  __device__ __forceinline__ static Vertex update(uint32_t, Acc acc, Vertex old_v, const Params&) {
    return acc < old_v ? acc : old_v;
  }
  __device__ __forceinline__ static bool better(Vertex cand, Vertex cur) { return cand < cur; }
  __device__ __forceinline__ static Vertex atomic_relax(Vertex* addr, Vertex cand) { return atomicMin(addr, cand); }
};

}  // namespace luxb
