// Round-2 microbenchmark (compiled here, to be run on a B200): can the hottest values be gathered from the
// DISTRIBUTED shared memory of a thread-block cluster instead of through the L1 tag stage?
//   variant A: plain ld.global.nc gathers (the 0.9 sectors/cycle/SM wall)
//   variant B: ids < HEAD are read from a per-CTA shared-memory copy (HEAD = 8 K values)
//   variant C: ids < HEAD are read with ld.shared::cluster from the owning CTA of a cluster of CL CTAs, each holding
//              HEAD / CL values (HEAD = 32 K * CL)
// Index stream: RMAT-27 source distribution remapped to popcount order (hot-packed), as in ubench_gather2.cu.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_dsmem_gather.bin scripts/ubench_dsmem_gather.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>
#include "../lux_b200/csrc/build.cuh"
namespace cg = cooperative_groups;
using namespace luxb;
namespace luxb { void set_error(const char*, ...) {} }

__global__ void popc_keys(uint32_t* keys, uint32_t* vals, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { keys[i] = __popc(i); vals[i] = i; }
}
__global__ void invert(const uint32_t* order, uint32_t* rank, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) rank[order[i]] = i;
}
__global__ void gen_idx(uint32_t* idx, uint64_t m, int scale, uint32_t n, const uint32_t* rank) {
  uint64_t sm = splitmix64(27);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t s, d;
    rmat_edge(sm, i, scale, n, s, d);
    idx[i] = rank[s];
  }
}

constexpr int U = 8;

__global__ void gather_plain(const uint32_t* __restrict__ idx, const float* __restrict__ x, uint64_t m, float* out) {
  float acc = 0.f;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < m; base += stride * U) {
    uint32_t id[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { uint64_t i = base + k * stride; id[k] = i < m ? __ldg(idx + i) : 0; }
#pragma unroll
    for (int k = 0; k < U; ++k) acc += __ldg(x + id[k]);
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int HEAD>
__global__ void gather_smem_head(const uint32_t* __restrict__ idx, const float* __restrict__ x, uint64_t m, float* out) {
  extern __shared__ float head[];
  for (int i = threadIdx.x; i < HEAD; i += blockDim.x) head[i] = x[i];
  __syncthreads();
  float acc = 0.f;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < m; base += stride * U) {
    uint32_t id[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { uint64_t i = base + k * stride; id[k] = i < m ? __ldg(idx + i) : 0; }
#pragma unroll
    for (int k = 0; k < U; ++k) acc += id[k] < HEAD ? head[id[k]] : __ldg(x + id[k]);
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int SLICE>  // values per CTA; cluster size comes from the launch attribute
__global__ void gather_dsmem_head(const uint32_t* __restrict__ idx, const float* __restrict__ x, uint64_t m, float* out) {
  extern __shared__ float slice[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cl = cluster.num_blocks(), me = cluster.block_rank();
  for (int i = threadIdx.x; i < SLICE; i += blockDim.x) slice[i] = x[(size_t)me * SLICE + i];
  cluster.sync();
  const uint32_t head = SLICE * cl;
  float acc = 0.f;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < m; base += stride * U) {
    uint32_t id[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { uint64_t i = base + k * stride; id[k] = i < m ? __ldg(idx + i) : 0; }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (id[k] < head) {
        const float* remote = cluster.map_shared_rank(slice, id[k] / SLICE);
        acc += remote[id[k] % SLICE];
      } else {
        acc += __ldg(x + id[k]);
      }
    }
  }
  cluster.sync();  // nobody may exit while its slice is still being read
  if (acc == 123.456f) out[0] = acc;
}

template <class F>
float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
  return best;
}

int main() {
  const int scale = 27;
  const uint32_t n = 1u << scale;
  const uint64_t m = 1ull << 29;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* x; uint32_t *idx, *keys, *keys2, *vals, *order, *rank; float* out;
  cudaMalloc(&x, (size_t)n * 4); cudaMemset(x, 0, (size_t)n * 4);
  cudaMalloc(&idx, m * 4); cudaMalloc(&out, 4);
  cudaMalloc(&keys, n * 4ull); cudaMalloc(&keys2, n * 4ull); cudaMalloc(&vals, n * 4ull); cudaMalloc(&order, n * 4ull); cudaMalloc(&rank, n * 4ull);
  popc_keys<<<sms * 8, 256>>>(keys, vals, n);
  size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, vals, order, (int)n, 0, 6);
  void* tmp; cudaMalloc(&tmp, tb); cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, vals, order, (int)n, 0, 6);
  invert<<<sms * 8, 256>>>(order, rank, n);
  gen_idx<<<sms * 16, 256>>>(idx, m, scale, n, rank);
  cudaDeviceSynchronize();
  float ms = timeit([&] { gather_plain<<<sms * 4, 256>>>(idx, x, m, out); });
  printf("plain ld.global.nc                         : %7.3f ms %6.1f Ggather/s\n", ms, m / ms / 1e6);
  {
    constexpr int HEAD = 8192;
    cudaFuncSetAttribute(gather_smem_head<HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD * 4);
    ms = timeit([&] { gather_smem_head<HEAD><<<sms * 4, 256, HEAD * 4>>>(idx, x, m, out); });
    printf("per-CTA smem head of %6d values          : %7.3f ms %6.1f Ggather/s\n", HEAD, ms, m / ms / 1e6);
  }
  for (int cl : {2, 4, 8}) {
    constexpr int SLICE = 32768;  // 128 KB per CTA, one CTA per SM
    cudaFuncSetAttribute(gather_dsmem_head<SLICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLICE * 4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((sms / cl) * cl);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = SLICE * 4;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = cl; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    ms = timeit([&] { cudaLaunchKernelEx(&cfg, gather_dsmem_head<SLICE>, (const uint32_t*)idx, (const float*)x, m, out); });
    cudaError_t e = cudaGetLastError();
    printf("DSMEM head, cluster %d x %6d values (%4d K) : %7.3f ms %6.1f Ggather/s  [%s]\n", cl, SLICE, cl * SLICE / 1024, ms,
           m / ms / 1e6, cudaGetErrorString(e));
  }
  return 0;
}
