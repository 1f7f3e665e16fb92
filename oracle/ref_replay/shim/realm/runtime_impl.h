// stand-in for realm/runtime_impl.h (see ../legion.h): the FB allocator the reference's init bodies use.
#pragma once
#include <cuda_runtime.h>
#include <chrono>
#include "legion.h"

namespace Realm {
struct MemoryImpl { virtual ~MemoryImpl() {} };
namespace Cuda {
// Realm's GPUFBMemory::alloc_bytes returns an OFFSET into the framebuffer pool; get_direct_ptr turns it into a
// pointer (pagerank_gpu.cu:265-277).  Here: every allocation is its own cudaMalloc (+64 KB pad: the reference's
// pr_kernel writes past newPrFb for the tail threads of the last block), offset = table index + 1.
struct GPUFBMemory : MemoryImpl {
  std::vector<void*> table;
  off_t alloc_bytes(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes + 65536) != cudaSuccess) return -1;
    cudaMemset(p, 0, bytes + 65536);
    table.push_back(p);
    return (off_t)table.size();
  }
  void free_bytes(off_t off, size_t) {
    if (off >= 1 && (size_t)off <= table.size() && table[off - 1]) { cudaFree(table[off - 1]); table[off - 1] = nullptr; }
  }
  void* get_direct_ptr(off_t off, size_t) { return table[off - 1]; }
  void release_all() { for (void*& p : table) if (p) { cudaFree(p); p = nullptr; } table.clear(); }
};
}  // namespace Cuda
struct RuntimeImpl {
  Cuda::GPUFBMemory fb;
  MemoryImpl* get_memory_impl(Legion::Memory) { return &fb; }
};
inline RuntimeImpl* get_runtime() { static RuntimeImpl rt; return &rt; }
struct Clock {
  static double current_time_in_microseconds() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
};
}  // namespace Realm
