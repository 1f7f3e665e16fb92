// Dev microbenchmark 2: gather rate vs array size and vs hot-vertex packing (relabel by hotness).
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>
#include "../lux_b200/csrc/build.cuh"
using namespace luxb;
namespace luxb { void set_error(const char*, ...) {} }

__global__ void gen_idx(uint32_t* idx, uint64_t m, int mode, int scale, uint32_t n, const uint32_t* rank) {
  uint64_t sm = splitmix64(27);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    if (mode == 0) idx[i] = (uint32_t)(splitmix64(i ^ 0x1234) % n);
    else { uint32_t s, d; rmat_edge(sm, i, scale, n, s, d); idx[i] = mode == 2 ? rank[s] : s; }
  }
}
__global__ void popc_keys(uint32_t* keys, uint32_t* vals, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { keys[i] = __popc(i); vals[i] = i; }
}
__global__ void invert(const uint32_t* order, uint32_t* rank, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) rank[order[i]] = i;
}
template <int U>
__global__ void gather_ldg(const uint32_t* __restrict__ idx, const float* __restrict__ x, uint64_t m, float* out) {
  float acc = 0.f;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < m; base += stride * U) {
    uint32_t id[U]; float v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { uint64_t i = base + k * stride; id[k] = i < m ? __ldg(idx + i) : 0; }
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = __ldg(x + id[k]);
#pragma unroll
    for (int k = 0; k < U; ++k) acc += v[k];
  }
  if (acc == 123.456f) out[0] = acc;
}
int main(int argc, char** argv) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint64_t m = 1ull << 29;
  uint32_t* idx; float* out; cudaMalloc(&idx, m * 4); cudaMalloc(&out, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int scale = 20; scale <= 28; ++scale) {
    uint32_t n = 1u << scale;
    float* x; cudaMalloc(&x, (size_t)n * 4); cudaMemset(x, 0, (size_t)n * 4);
    uint32_t *keys, *keys2, *vals, *order, *rank;
    cudaMalloc(&keys, n * 4ull); cudaMalloc(&keys2, n * 4ull); cudaMalloc(&vals, n * 4ull); cudaMalloc(&order, n * 4ull); cudaMalloc(&rank, n * 4ull);
    popc_keys<<<sms * 8, 256>>>(keys, vals, n);
    size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, vals, order, (int)n, 0, 6);
    void* tmp; cudaMalloc(&tmp, tb); cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, vals, order, (int)n, 0, 6);
    invert<<<sms * 8, 256>>>(order, rank, n);
    for (int mode = 0; mode < 3; ++mode) {
      gen_idx<<<sms * 16, 256>>>(idx, m, mode, scale, n, rank);
      gather_ldg<8><<<sms * 4, 256>>>(idx, x, m, out); cudaDeviceSynchronize();
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) { cudaEventRecord(a); gather_ldg<8><<<sms * 4, 256>>>(idx, x, m, out); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
      printf("n=2^%d (%5.0f MB) %-14s: %7.3f ms %6.1f Ggather/s\n", scale, n * 4.0 / 1e6, mode == 0 ? "uniform" : mode == 1 ? "rmat" : "rmat-hot-packed", best, m / best / 1e6);
    }
    cudaFree(x); cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(order); cudaFree(rank); cudaFree(tmp);
  }
  return 0;
}
