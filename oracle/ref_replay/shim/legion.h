// Minimal stand-in for legion.h — just enough surface for /root/reference/core/graph.h and the task BODIES in
// <app>/<app>_gpu.cu to compile unmodified and be called directly (no runtime, no scheduling, no data movement).
// TEST INFRASTRUCTURE (oracle/ref_replay): lets the reference's own CUDA kernels run on the GPU box so that their
// outputs can pin the oracle and their time can stand beside ours.  Not part of the product.
#pragma once
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace Realm {
template <typename FT, int N, typename T>
struct AffineAccessor {
  template <class R>
  bool is_dense_arbitrary(const R&) const { return true; }
};
}  // namespace Realm

namespace Legion {
typedef long long coord_t;
enum PrivilegeMode { READ_ONLY, READ_WRITE, WRITE_ONLY };
typedef unsigned FieldID;

template <int N, typename T = coord_t>
struct Point {
  T x[N];
  Point() { for (int i = 0; i < N; i++) x[i] = 0; }
  Point(T v) { for (int i = 0; i < N; i++) x[i] = v; }
  T& operator[](int i) { return x[i]; }
  const T& operator[](int i) const { return x[i]; }
};
template <int N, typename T = coord_t>
struct Rect {
  Point<N, T> lo, hi;
  Rect() {}
  Rect(Point<N, T> l, Point<N, T> h) : lo(l), hi(h) {}
  bool operator==(const Rect& o) const { return lo[0] == o.lo[0] && hi[0] == o.hi[0]; }
};
struct Domain {
  coord_t lo, hi;
  template <int N, typename T>
  operator Rect<N, T>() const { return Rect<N, T>(Point<N, T>((T)lo), Point<N, T>((T)hi)); }
};
struct IndexSpace { coord_t lo, hi; };
template <int N, typename T = coord_t>
struct IndexSpaceT : IndexSpace {};
struct LogicalRegion {
  coord_t lo, hi;
  IndexSpace get_index_space() const { IndexSpace s; s.lo = lo; s.hi = hi; return s; }
};
struct LogicalPartition { int unused; };
struct Memory {
  enum Kind { GPU_FB_MEM, Z_COPY_MEM };
  Kind k;
  int id;
  Kind kind() const { return k; }
  bool operator<(const Memory& o) const { return id < o.id; }
};
struct RegionRequirement { LogicalRegion region; };
struct Task {
  void* args;
  void* local_args;
  std::vector<RegionRequirement> regions;
};
struct PhysicalRegion {
  void* base;       // address of the element at index lo
  coord_t lo, hi;
  Memory mem;
  void get_memories(std::set<Memory>& s) const { s.insert(mem); }
};
struct ContextImpl {};
typedef ContextImpl* Context;
struct Runtime {
  Domain get_index_space_domain(Context, IndexSpace is) { Domain d; d.lo = is.lo; d.hi = is.hi; return d; }
};
typedef Runtime HighLevelRuntime;

template <PrivilegeMode M, typename FT, int N, typename T, typename A>
struct FieldAccessor {
  A accessor;
  FT* base;
  coord_t lo;
  FieldAccessor(const PhysicalRegion& r, FieldID) : base((FT*)r.base), lo(r.lo) {}
  FT* ptr(const Rect<N, T>& r) const { return base + (r.lo[0] - lo); }
  FT* ptr(const Point<N, T>& p) const { return base + (p[0] - lo); }
};

struct ArgumentMap {};
struct TaskArgument {};
struct IndexLauncher {};
struct TaskLauncher {};
}  // namespace Legion
