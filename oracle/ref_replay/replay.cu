// replay.cu — calls the reference's OWN task bodies and CUDA kernels (compiled from /root/reference at build time,
// nothing copied) on fabricated Task / PhysicalRegion objects, one partition, one GPU.
// Built three times by build.py with -DREF_APP_{PAGERANK,COMPONENTS,SSSP} and -I<reference>/<app>:
//   REF_CU expands to the absolute path of <app>/<app>_gpu.cu.
// Zero-copy regions (MAP_TO_ZC_MEMORY, core/graph.h:34) are emulated with mapped pinned host memory, the FB pool
// with cudaMalloc (shim/realm/runtime_impl.h).  TEST INFRASTRUCTURE — never linked into libluxb.
#include REF_CU

#include <chrono>

using namespace Legion;

namespace {
template <class T>
T* zc_alloc(size_t n) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, (n ? n : 1) * sizeof(T) + 4096, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
  memset(p, 0, (n ? n : 1) * sizeof(T) + 4096);
  return (T*)p;
}
template <class T>
T* fb_alloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, (n ? n : 1) * sizeof(T) + 65536) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, (n ? n : 1) * sizeof(T) + 65536);
  return (T*)p;
}
PhysicalRegion region(void* base, coord_t lo, coord_t hi, bool fb) {
  PhysicalRegion r;
  r.base = base; r.lo = lo; r.hi = hi;
  r.mem.k = fb ? Memory::GPU_FB_MEM : Memory::Z_COPY_MEM;
  r.mem.id = fb ? 1 : 2;
  return r;
}
void add(Task& t, std::vector<PhysicalRegion>& v, const PhysicalRegion& r) {
  RegionRequirement rr;
  rr.region.lo = r.lo; rr.region.hi = r.hi;
  t.regions.push_back(rr);
  v.push_back(r);
}
Graph* make_graph(V_ID nv, E_ID ne) {
  Graph* g = (Graph*)calloc(1, sizeof(Graph));  // Graph's only constructor needs Legion; its data members are PODs
  g->numParts = 1; g->nv = nv; g->ne = ne;
  g->rowLeft[0] = 0; g->rowRight[0] = nv - 1;
  g->verbose = false;
  return g;
}
}  // namespace

#if defined(REF_APP_PAGERANK)
// returns 0; out[nv] = dist_lr[ni % 2] as the reference would hold it; loop_ms = wall time of the ni PullAppTasks
extern "C" int ref_pagerank(uint32_t nv, uint64_t ne, const uint64_t* row_end, const uint32_t* src, int ni, float* out,
                            double* loop_ms) {
  Runtime rt;
  Graph* graph = make_graph(nv, ne);
  E_ID* raw_rows = zc_alloc<E_ID>(nv);
  V_ID* raw_cols = zc_alloc<V_ID>(ne);
  V_ID* degrees = zc_alloc<V_ID>(nv);
  Vertex* dist[2] = {zc_alloc<Vertex>(nv), zc_alloc<Vertex>(nv)};
  memcpy(raw_rows, row_end, sizeof(E_ID) * nv);
  memcpy(raw_cols, src, sizeof(V_ID) * ne);
  for (E_ID e = 0; e < ne; e++) degrees[src[e]]++;  // pull_scan_task_impl, pull_model.inl:333-343 (CPU task in core)
  NodeStruct* row_ptrs = fb_alloc<NodeStruct>(nv);
  V_ID* in_vtxs = fb_alloc<V_ID>(ne);
  EdgeStruct* col_idxs = fb_alloc<EdgeStruct>(ne);
  // PullInitTask region contract: pull_model.inl:347-421
  Task init;
  std::vector<PhysicalRegion> ir;
  init.args = graph; init.local_args = nullptr;
  add(init, ir, region(row_ptrs, 0, nv - 1, true));
  add(init, ir, region(in_vtxs, 0, ne - 1, true));
  add(init, ir, region(col_idxs, 0, ne - 1, true));
  add(init, ir, region(dist[0], 0, nv - 1, false));
  add(init, ir, region(raw_rows, 0, nv - 1, false));
  add(init, ir, region(raw_cols, 0, ne - 1, false));
  add(init, ir, region(degrees, 0, nv - 1, false));
  GraphPiece piece = pull_init_task_impl(&init, ir, nullptr, &rt);
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < ni; i++) {  // pagerank.cc:109-113; PullAppTask regions: pull_model.inl:423-470
    Task app;
    std::vector<PhysicalRegion> ar;
    app.args = graph; app.local_args = &piece;
    add(app, ar, region(row_ptrs, 0, nv - 1, true));
    add(app, ar, region(in_vtxs, 0, ne - 1, true));
    add(app, ar, region(col_idxs, 0, ne - 1, true));
    add(app, ar, region(dist[i % 2], 0, nv - 1, false));
    add(app, ar, region(dist[(i + 1) % 2], 0, nv - 1, false));
    pull_app_task_impl(&app, ar, nullptr, &rt);
  }
  cudaDeviceSynchronize();
  *loop_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  memcpy(out, dist[ni % 2], sizeof(Vertex) * nv);
  cudaFreeHost(raw_rows); cudaFreeHost(raw_cols); cudaFreeHost(degrees); cudaFreeHost(dist[0]); cudaFreeHost(dist[1]);
  cudaFree(row_ptrs); cudaFree(in_vtxs); cudaFree(col_idxs);
  Realm::get_runtime()->fb.release_all();
  free(graph);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
#endif

#if defined(REF_APP_COMPONENTS) || defined(REF_APP_SSSP)
// labels[nv] = final dist_lr; active[it] = numNodes returned by PushAppTask it; returns #iterations (incl. the final 0)
extern "C" int ref_labels(uint32_t nv, uint64_t ne, const uint64_t* row_end, const uint32_t* src, uint32_t start,
                          uint32_t* labels, uint32_t* active, int max_iters, double* loop_ms, uint32_t* check_mistakes) {
  Runtime rt;
  Graph* graph = make_graph(nv, ne);
  graph->startVtx = start;
  V_ID slots = (nv - 1) / SPARSE_THRESHOLD + 100;  // push_model.inl:393
  graph->frontierSize = sizeof(FrontierHeader) + slots * sizeof(V_ID);
  graph->fqLeft[0] = 0; graph->fqRight[0] = graph->frontierSize - 1;
  E_ID* raw_rows = zc_alloc<E_ID>(nv);
  V_ID* raw_cols = zc_alloc<V_ID>(ne);
  Vertex* dist[2] = {zc_alloc<Vertex>(nv), zc_alloc<Vertex>(nv)};
  char* fq[2] = {zc_alloc<char>(graph->frontierSize), zc_alloc<char>(graph->frontierSize)};
  memcpy(raw_rows, row_end, sizeof(E_ID) * nv);
  memcpy(raw_cols, src, sizeof(V_ID) * ne);
  NodeStruct* pull_row = fb_alloc<NodeStruct>(nv);
  EdgeStruct2* pull_col = fb_alloc<EdgeStruct2>(ne);
  NodeStruct* push_row = fb_alloc<NodeStruct>(nv);
  EdgeStruct* push_col = fb_alloc<EdgeStruct>(ne);
  coord_t fhi = (coord_t)graph->frontierSize - 1;
  Task init;
  std::vector<PhysicalRegion> ir;  // PushInitTask contract: push_model.inl:123-194
  init.args = graph; init.local_args = nullptr;
  add(init, ir, region(pull_row, 0, nv - 1, true));
  add(init, ir, region(pull_col, 0, ne - 1, true));
  add(init, ir, region(push_row, 0, nv - 1, true));
  add(init, ir, region(push_col, 0, ne - 1, true));
  add(init, ir, region(fq[0], 0, fhi, false));
  add(init, ir, region(dist[0], 0, nv - 1, false));
  add(init, ir, region(raw_rows, 0, nv - 1, false));
  add(init, ir, region(raw_cols, 0, ne - 1, false));
  GraphPiece piece = push_init_task_impl(&init, ir, nullptr, &rt);
  cudaDeviceSynchronize();
  int it = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (; it < max_iters;) {  // components.cc:113-127 without the 4-deep window; regions: push_model.inl:196-265
    Task app;
    std::vector<PhysicalRegion> ar;
    app.args = graph; app.local_args = &piece;
    add(app, ar, region(pull_row, 0, nv - 1, true));
    add(app, ar, region(pull_col, 0, ne - 1, true));
    add(app, ar, region(push_row, 0, nv - 1, true));
    add(app, ar, region(push_col, 0, ne - 1, true));
    add(app, ar, region(fq[it % 2], 0, fhi, false));
    add(app, ar, region(fq[(it + 1) % 2], 0, fhi, false));
    add(app, ar, region(dist[it % 2], 0, nv - 1, false));
    add(app, ar, region(dist[(it + 1) % 2], 0, nv - 1, false));
    V_ID n_active = push_app_task_impl(&app, ar, nullptr, &rt);
    active[it] = n_active;
    ++it;
    if (n_active == 0) break;
  }
  cudaDeviceSynchronize();
  *loop_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  memcpy(labels, dist[it % 2], sizeof(Vertex) * nv);
  // CheckTask (components_gpu.cu:794-837): prints [PASS]/[FAIL]; we also recount on the host for the caller
  {
    Task chk;
    std::vector<PhysicalRegion> cr;
    chk.args = graph; chk.local_args = &piece;
    add(chk, cr, region(pull_row, 0, nv - 1, true));
    add(chk, cr, region(pull_col, 0, ne - 1, true));
    add(chk, cr, region(dist[it % 2], 0, nv - 1, false));
    check_task_impl(&chk, cr, nullptr, &rt);
  }
  uint32_t bad = 0;
  for (V_ID v = 0; v < nv; v++)
    for (E_ID e = (v == 0 ? 0 : row_end[v - 1]); e < row_end[v]; e++) {
#if defined(REF_APP_COMPONENTS)
      bad += labels[v] < labels[src[e]];
#else
      bad += (labels[src[e]] != nv) && (labels[v] > labels[src[e]] + 1);
#endif
    }
  *check_mistakes = bad;
  cudaFreeHost(raw_rows); cudaFreeHost(raw_cols); cudaFreeHost(dist[0]); cudaFreeHost(dist[1]); cudaFreeHost(fq[0]); cudaFreeHost(fq[1]);
  cudaFree(pull_row); cudaFree(pull_col); cudaFree(push_row); cudaFree(push_col);
  Realm::get_runtime()->fb.release_all();
  free(graph);
  return it;
}
#endif
