// cf.cuh — collaborative filtering (SGD on K = 20 latent factors), pull model.  Replaces cf_kernel
// (colfilter_gpu.cu:32-104) with the INTENDED math: the reference kernel has a shared-memory data
// race, a wrong transpose index, a divergent barrier and a local/global index mix-up which
// are not reproduced.
//   err(u,v,w) = w - <x_u, x_v>;  acc_v = sum_in-edges err * x_u;  x_v' = x_v + GAMMA * (acc_v - LAMBDA * x_v)
// Work decomposition: every destination vertex's in-edge list is cut into chunks of kCfChunk edges; ONE WARP
// owns one chunk.  Inside a warp, 4 groups of 8 lanes each take one edge at a time; lanes 0-4 of a group hold
// one float4 (4 of the 20 factors) so a source vector is fetched with five coalesced 128-bit loads per edge.
// Chunk partials are combined in fixed (ascending) order in fp64 by cf_update_kernel -> deterministic, race-free.
#pragma once
#include "common.cuh"

namespace luxb {

constexpr int kCfK = LUXB_CF_K;
constexpr int kCfChunk = 256;
#ifndef LUXB_CF_UNROLL
#define LUXB_CF_UNROLL 4
#endif
constexpr int kCfUnroll = LUXB_CF_UNROLL;  // 32 edges per load round must be a multiple of 4 * kCfUnroll
constexpr float kCfLambda = 0.001f;       // LAMBDA, col_filter/app.h:26
constexpr float kCfGamma = 0.00000035f;   // GAMMA,  col_filter/app.h:27

__global__ void cf_chunk_count_kernel(const uint64_t* __restrict__ row_end_rel, uint32_t n_part, uint32_t* __restrict__ cnt) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_part; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t d = row_end_rel[i] - (i == 0 ? 0 : row_end_rel[i - 1]);
    uint64_t c = (d + kCfChunk - 1) / kCfChunk;
    cnt[i] = (uint32_t)(c ? c : 1);
  }
}

// chunk_first: exclusive scan of counts (n_part + 1 entries) -> chunk_vtx[c] = owning local vertex
__global__ void cf_chunk_fill_kernel(const uint32_t* __restrict__ chunk_first, uint32_t n_part, uint32_t* __restrict__ chunk_vtx) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_part; i += (uint64_t)gridDim.x * blockDim.x)
    for (uint32_t c = chunk_first[i]; c < chunk_first[i + 1]; ++c) chunk_vtx[c] = (uint32_t)i;
}

struct CfArgs {
  const uint64_t* row_end;     // relative, padded
  const uint32_t* src;
  const int32_t* weight;
  const uint32_t* chunk_first; // [n_part + 1]
  const uint32_t* chunk_vtx;   // [n_chunks]
  uint32_t n_part, n_chunks, row_left;
  const float* x_old;          // [nv * 20] replica
  float* partial;              // [n_chunks * 20]
  float* out;                  // [n_part * 20] this partition's new vectors (local index)
  int n_peers;
  float* peer_out[LUXB_MAX_PEERS];
};

__global__ void __launch_bounds__(256) cf_chunk_kernel(const __grid_constant__ CfArgs a) {
  const unsigned lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
  const bool holder = sub < 5;
  uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < a.n_chunks; c += warps_total) {
    uint32_t v = __ldg(a.chunk_vtx + c);
    uint32_t k = c - __ldg(a.chunk_first + v);
    uint64_t vb = v == 0 ? 0 : __ldg(a.row_end + v - 1), ve = __ldg(a.row_end + v);
    uint64_t b = vb + (uint64_t)k * kCfChunk;
    uint64_t e = b + kCfChunk < ve ? b + kCfChunk : ve;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (holder) xv = __ldg(reinterpret_cast<const float4*>(a.x_old + (size_t)(a.row_left + v) * kCfK) + sub);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint64_t base = b; base < e; base += 32) {
      uint64_t me = base + lane;
      uint32_t s = 0;
      int32_t w = 0;
      if (me < e) { s = __ldg(a.src + me); w = __ldg(a.weight + me); }
      uint32_t n = e - base < 32 ? (uint32_t)(e - base) : 32u;
      // kCfUnroll edges per group in flight: all source vectors are requested before any is used (the kernel is bound
      // by L2 gather latency — ncu long_scoreboard — so loads in flight per lane are what buys time)
      for (uint32_t q = 0; q < n; q += 4 * kCfUnroll) {
        float4 xu[kCfUnroll];
        int32_t wu[kCfUnroll];
        bool valid[kCfUnroll];
#pragma unroll
        for (int u = 0; u < kCfUnroll; ++u) {
          const uint32_t idx = q + 4 * u + grp;
          const uint32_t su = __shfl_sync(0xffffffffu, s, idx & 31);
          wu[u] = __shfl_sync(0xffffffffu, w, idx & 31);
          valid[u] = idx < n;
          xu[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid[u] && holder) xu[u] = __ldg(reinterpret_cast<const float4*>(a.x_old + (size_t)su * kCfK) + sub);
        }
#pragma unroll
        for (int u = 0; u < kCfUnroll; ++u) {  // ascending edge order inside the group: same numerics for any unroll
          float dot = xu[u].x * xv.x;
          dot = __fmaf_rn(xu[u].y, xv.y, dot);
          dot = __fmaf_rn(xu[u].z, xv.z, dot);
          dot = __fmaf_rn(xu[u].w, xv.w, dot);
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
          dot += __shfl_xor_sync(0xffffffffu, dot, 4);
          const float err = valid[u] ? (float)wu[u] - dot : 0.f;
          acc.x = __fmaf_rn(err, xu[u].x, acc.x);
          acc.y = __fmaf_rn(err, xu[u].y, acc.y);
          acc.z = __fmaf_rn(err, xu[u].z, acc.z);
          acc.w = __fmaf_rn(err, xu[u].w, acc.w);
        }
      }
    }
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
    }
    if (grp == 0 && holder) reinterpret_cast<float4*>(a.partial + (size_t)c * kCfK)[sub] = acc;
  }
}

// one thread per (vertex, factor): ordered fp64 sum of the vertex's chunk partials, then the SGD update
__global__ void cf_update_kernel(const __grid_constant__ CfArgs a) {
  uint64_t total = (uint64_t)a.n_part * kCfK;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = (uint32_t)(t / kCfK), f = (uint32_t)(t % kCfK);
    double acc = 0.0;
    for (uint32_t c = a.chunk_first[v]; c < a.chunk_first[v + 1]; ++c) acc += (double)a.partial[(size_t)c * kCfK + f];
    float xv = a.x_old[(size_t)(a.row_left + v) * kCfK + f];
    float nx = xv + kCfGamma * ((float)acc - kCfLambda * xv);  // colfilter_gpu.cu:99-100
    a.out[t] = nx;
    for (int p = 0; p < a.n_peers; ++p) a.peer_out[p][t] = nx;
  }
}

__global__ void cf_init_kernel(float* x, uint64_t n) {
  float value = sqrtf(__fdiv_rn(1.0f, (float)kCfK));  // colfilter_gpu.cu:260
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) x[i] = value;
}

}  // namespace luxb
