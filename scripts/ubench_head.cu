// Round-2 microbenchmark: where can the hottest vertex values live so that their gathers do not pass the L1 tag stage?
// Index stream: 2^29 draws of the RMAT-27 source distribution, remapped to popcount order (= hot-packed rank), the
// same stream ubench_gather2.cu uses.  Every variant gathers x[id] and sums it; they differ in where ids < HEAD are read:
//   plain   : ld.global.nc (the ~0.9 sectors/cycle/SM wall)
//   smem    : per-CTA shared-memory copy of the first HEAD values                  (HEAD = 8K .. 48K)
//   dsmem   : ld.shared::cluster from the owning CTA of a CL-CTA cluster, SLICE values per CTA (HEAD = CL * SLICE)
//   pure_*  : every id folded into the head (isolates the shared / distributed-shared gather rate itself)
// Also reports how many distinct SMs a cluster launch of grid = #SMs actually occupies.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_head.bin scripts/ubench_head.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>
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
__global__ void fold_idx(const uint32_t* in, uint32_t* out, uint64_t m, uint32_t mod) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i] % mod;
}
__global__ void count_below(const uint32_t* in, uint64_t m, uint32_t thr, unsigned long long* cnt) {
  unsigned long long c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) c += in[i] < thr;
  atomicAdd(cnt, c);
}

constexpr int U = 8;

__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

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

__global__ void gather_smem_head(const uint32_t* __restrict__ idx, const float* __restrict__ x, uint64_t m, float* out, uint32_t head) {
  extern __shared__ float tab[];
  for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) tab[i] = x[i];
  __syncthreads();
  float acc = 0.f;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < m; base += stride * U) {
    uint32_t id[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { uint64_t i = base + k * stride; id[k] = i < m ? __ldg(idx + i) : 0; }
    float v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) if (id[k] >= head) v[k] = __ldg(x + id[k]);
#pragma unroll
    for (int k = 0; k < U; ++k) if (id[k] < head) v[k] = tab[id[k]];
#pragma unroll
    for (int k = 0; k < U; ++k) acc += v[k];
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int LOG_SLICE>  // values per CTA = 1 << LOG_SLICE; cluster size comes from the launch attribute
__global__ void gather_dsmem_head(const uint32_t* __restrict__ idx, const float* __restrict__ x, uint64_t m, float* out, uint32_t* smids) {
  extern __shared__ float tab[];
  constexpr uint32_t SLICE = 1u << LOG_SLICE;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cl = cluster.num_blocks(), me = cluster.block_rank();
  for (uint32_t i = threadIdx.x; i < SLICE; i += blockDim.x) tab[i] = x[(size_t)me * SLICE + i];
  if (threadIdx.x == 0 && smids) smids[blockIdx.x] = smid();
  cluster.sync();
  const uint32_t head = SLICE * cl;
  const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab);
  float acc = 0.f;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < m; base += stride * U) {
    uint32_t id[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { uint64_t i = base + k * stride; id[k] = i < m ? __ldg(idx + i) : 0; }
    float v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) if (id[k] >= head) v[k] = __ldg(x + id[k]);
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (id[k] < head) {
        uint32_t ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(tab_s + ((id[k] & (SLICE - 1)) << 2)), "r"(id[k] >> LOG_SLICE));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v[k]) : "r"(ra));
      }
#pragma unroll
    for (int k = 0; k < U; ++k) acc += v[k];
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
  cudaEventDestroy(a); cudaEventDestroy(b);
  return best;
}

static const uint32_t* g_idx; static const float* g_x; static uint64_t g_m; static float* g_out; static uint32_t* g_smids; static int g_sms;

template <int LOG_SLICE>
void run_dsmem(int cl, int threads, const uint32_t* idx, const char* tag) {
  constexpr uint32_t SLICE = 1u << LOG_SLICE;
  auto kern = gather_dsmem_head<LOG_SLICE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SLICE * 4);
  if (cl > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((g_sms / cl) * cl);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = SLICE * 4;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cl; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int max_clusters = -1;
  cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  cudaMemset(g_smids, 0xFF, 4096 * 4);
  float ms = timeit([&] { cudaLaunchKernelEx(&cfg, kern, idx, g_x, g_m, g_out, g_smids); });
  cudaError_t e = cudaGetLastError();
  std::vector<uint32_t> h(cfg.gridDim.x);
  cudaMemcpy(h.data(), g_smids, h.size() * 4, cudaMemcpyDeviceToHost);
  std::set<uint32_t> distinct(h.begin(), h.end());
  printf("dsmem%-6s cluster %2d x %6u values (%5u K head) %4d thr grid %3d: %7.3f ms %6.1f Ggather/s  SMs used %3zu  maxActiveClusters %d [%s]\n",
         tag, cl, SLICE, cl * SLICE / 1024, threads, cfg.gridDim.x, ms, g_m / ms / 1e6, distinct.size(), max_clusters, cudaGetErrorString(e));
  fflush(stdout);
}

int main() {
  const int scale = 27;
  const uint32_t n = 1u << scale;
  const uint64_t m = 1ull << 29;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  g_sms = sms;
  float* x; uint32_t *idx, *idx2, *keys, *keys2, *vals, *order, *rank; float* out; uint32_t* smids; unsigned long long* cnt;
  cudaMalloc(&x, (size_t)n * 4); cudaMemset(x, 0, (size_t)n * 4);
  cudaMalloc(&idx, m * 4); cudaMalloc(&idx2, m * 4); cudaMalloc(&out, 4); cudaMalloc(&smids, 4096 * 4); cudaMalloc(&cnt, 8);
  cudaMalloc(&keys, n * 4ull); cudaMalloc(&keys2, n * 4ull); cudaMalloc(&vals, n * 4ull); cudaMalloc(&order, n * 4ull); cudaMalloc(&rank, n * 4ull);
  popc_keys<<<sms * 8, 256>>>(keys, vals, n);
  size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, vals, order, (int)n, 0, 6);
  void* tmp; cudaMalloc(&tmp, tb); cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, vals, order, (int)n, 0, 6);
  invert<<<sms * 8, 256>>>(order, rank, n);
  gen_idx<<<sms * 16, 256>>>(idx, m, scale, n, rank);
  cudaDeviceSynchronize();
  g_idx = idx; g_x = x; g_m = m; g_out = out; g_smids = smids;
  for (uint32_t thr : {8192u, 32768u, 49152u, 131072u, 262144u, 524288u, 1048576u, 4194304u, 16777216u}) {
    cudaMemset(cnt, 0, 8);
    count_below<<<sms * 8, 256>>>(idx, m, thr, cnt);
    unsigned long long c; cudaMemcpy(&c, cnt, 8, cudaMemcpyDeviceToHost);
    printf("coverage: ids < %8u : %.4f of the gathers\n", thr, (double)c / m);
  }
  float ms;
  for (int ctas : {4, 8}) {
    ms = timeit([&] { gather_plain<<<sms * ctas, 256>>>(idx, x, m, out); });
    printf("plain ld.global.nc, %d CTAs/SM x 256                    : %7.3f ms %6.1f Ggather/s\n", ctas, ms, m / ms / 1e6);
  }
  // plain gathers while a big shared-memory carve-out shrinks L1 (the kernel allocates but does not use the table)
  for (uint32_t head : {0u, 8192u, 16384u, 32768u, 49152u}) {
    for (int threads : {512, 1024}) {
      size_t smem = (size_t)head * 4;
      int per_sm = smem ? (int)std::min<size_t>(2048 / threads, (220 * 1024) / (smem + 1024)) : 2048 / threads;
      if (per_sm < 1) continue;
      cudaFuncSetAttribute(gather_smem_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(gather_smem_head, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)std::min<size_t>(100, (per_sm * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)));
      ms = timeit([&] { gather_smem_head<<<sms * per_sm, threads, smem>>>(idx, x, m, out, head); });
      printf("smem head %6u values, %d CTAs/SM x %4d thr            : %7.3f ms %6.1f Ggather/s [%s]\n", head, per_sm, threads, ms,
             m / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      fflush(stdout);
    }
  }
  // pure shared-memory gather rate: all ids folded into the table
  fold_idx<<<sms * 8, 256>>>(idx, idx2, m, 32768);
  cudaFuncSetAttribute(gather_smem_head, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
  ms = timeit([&] { gather_smem_head<<<sms, 1024, 32768 * 4>>>(idx2, x, m, out, 32768); });
  printf("pure smem gathers (ids %% 32768), 1 CTA/SM x 1024         : %7.3f ms %6.1f Ggather/s\n", ms, m / ms / 1e6);

  for (int cl : {2, 4, 8, 16}) run_dsmem<15>(cl, 1024, idx, "");
  for (int cl : {2, 4, 8, 16}) run_dsmem<14>(cl, 1024, idx, "");
  for (int cl : {4, 8}) run_dsmem<15>(cl, 512, idx, "");
  // pure distributed-shared gather rate
  for (int cl : {2, 8}) {
    fold_idx<<<sms * 8, 256>>>(idx, idx2, m, 32768u * cl);
    run_dsmem<15>(cl, 1024, idx2, "-pure");
  }
  return 0;
}
