// runtime.cuh — the per-rank host runtime state (replaces Graph + GraphPiece, core/graph.h:54-98, and the
// placement/ownership role of LuxMapper + Realm's FB allocator).  One luxb_graph = one rank = one GPU.
#pragma once
#include <vector>
#include "comm.h"
#include "common.cuh"
#include "panel.cuh"
#include "seg.cuh"
#include "push.cuh"

// one CSC swept by a merge-path tile kernel, with its tile table and fix-up scratch
struct PullLayout {
  uint64_t* d_row_end = nullptr;    // [n_vtx + 4] (may be released once the tile table exists)
  uint32_t* d_row_end32 = nullptr;  // [n_vtx + 8]
  void* d_src = nullptr;            // u32 gather ids (main) or u16 block-local offsets (panel)
  uint32_t* d_tile_v = nullptr;
  uint32_t n_vtx = 0;
  uint64_t e_cnt = 0;
  uint32_t n_tiles = 0;
  void* d_head = nullptr;
  void* d_tail = nullptr;
  void* d_carry = nullptr;
  uint32_t* d_carry_flag = nullptr;
  void* d_block_agg = nullptr;
  uint32_t* d_block_flag = nullptr;
  uint32_t n_fix_blocks = 0;
  unsigned long long* d_chain = nullptr;  // fused fix-up: [2 * n_fix_blocks] values then [2 * n_fix_blocks] status words
  uint32_t chain_epoch = 0;
  // flagged stream (seg.cuh): d_src holds the words, d_tile_v the heads before each piece, n_tiles the pieces
  uint32_t* d_close = nullptr;      // [1 + heads] vertex completed by each head
  uint32_t* d_empty = nullptr;      // vertices without edges in this stream that are not hubs
  uint32_t n_empty = 0;
  uint32_t* d_empty_hub = nullptr;  // ... that are hubs (all their edges moved to the panel)
  uint32_t n_empty_hub = 0;
  uint32_t n_stages = 0;
  uint64_t n_words = 0;
};

// dev aid: LUXB_PHASE_TIMING=1 prints the mean device time of each phase of a PageRank iteration at luxb_close
struct PhaseTimer {
  bool on = false, per_call = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> tag;
  double sum[12] = {0};
  long cnt = 0;
};

struct luxb_graph {
  luxb_config cfg{};
  uint32_t nv = 0;
  uint64_t ne = 0;
  int P = 1;
  int parts_found = 0;  // partitions the reference's greedy scan produced (== P when the reference would accept)
  bool weighted = false;

  // global partition table (Graph::rowLeft/rowRight, core/graph.h:62)
  // the reference's split as reported by luxb_partition_bounds; rl / np / cl below are the WORK split in use (the same
  // unless cfg.balanced_split)
  uint32_t ref_rl[LUXB_MAX_PARTS]{};
  uint32_t ref_np[LUXB_MAX_PARTS]{};
  uint64_t ref_cl[LUXB_MAX_PARTS]{};
  uint32_t rl[LUXB_MAX_PARTS]{};
  uint32_t np[LUXB_MAX_PARTS]{};
  uint64_t cl[LUXB_MAX_PARTS]{};
  uint32_t cap[LUXB_MAX_PARTS]{};      // frontier queue capacity per partition (push_model.inl:393)
  uint64_t slot_off[LUXB_MAX_PARTS]{}; // our slot offsets inside fq buffers
  uint64_t slot_bytes[LUXB_MAX_PARTS]{};
  uint64_t fq_total = 0;

  // this rank's slice
  uint32_t row_left = 0, n_part = 0;
  uint64_t col_left = 0, e_part = 0;
  uint64_t* d_row_end = nullptr;  // [n_part + 4] relative end offsets + sentinels
  uint32_t* d_row_end32 = nullptr;  // [n_part + 8] low words of d_row_end (streamed by the tile kernel)
  int pull_shape = 1;
  uint32_t* d_src = nullptr;      // [e_part + 8]
  int32_t* d_weight = nullptr;    // [e_part + 8] (col_filter)
  uint32_t* d_tile_v = nullptr;
  uint32_t n_tiles = 0;
  void* d_head = nullptr;
  void* d_tail = nullptr;
  void* d_carry = nullptr;        // fix-up scratch (per tile / per block of tiles)
  uint32_t* d_carry_flag = nullptr;
  void* d_block_agg = nullptr;
  uint32_t* d_block_flag = nullptr;
  uint32_t n_fix_blocks = 0;

  cudaStream_t stream = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  int num_sms = 0;

  // app state
  bool inited = false;
  uint32_t* d_deg = nullptr;   // PageRank: global out-degrees
  void* d_val[2] = {nullptr, nullptr};  // replicas of the vertex values (labels: only [0])
  int cur = 0;
  size_t vbytes = 4;           // bytes per vertex value
  // hot-packed gather layout (PageRank): hot copies live in d_hot, natural-order values in d_val[0/1]
  uint32_t hot_n = 0;
  void* d_hot = nullptr;             // [hot_n] hot copies (single buffer: refreshed in place after every iteration)
  uint32_t hot_off[LUXB_MAX_PARTS + 1]{};  // hot slots owned by partition p: [hot_off[p], hot_off[p+1])
  uint32_t* d_hot_order = nullptr;   // [hot_n] vertex id held by each hot slot (descending out-degree)
  uint32_t* d_src_gather = nullptr;  // [e_part + 8] source ids rewritten as indices into Z
  // packed exchange (PageRank, nranks > 1): transfer arrays XT[2] = [hot by owner (hot_n) | cold-active by id (cold_n)]
  bool packed = false;
  uint32_t cold_n = 0;                       // vertices with 0 < out-degree < tau (all partitions)
  uint32_t cold_off[LUXB_MAX_PARTS + 1]{};   // cold-active vertices owned by partition p: [cold_off[p], cold_off[p+1])
  uint32_t* d_zperm = nullptr;               // [hot_n] global hot rank of the k-th entry of XT's hot part
  uint32_t* d_pack_list = nullptr;           // local indices of this rank's [hot | cold-active] vertices in transfer order
  float* d_xt[2] = {nullptr, nullptr};
  int cur_xt = 0;
  uint64_t xt_hot_chunk = 0, xt_cold_chunk = 0;  // equal chunks (elements) of the two balanced all-gathers; XT = [P hot chunks | P cold chunks]
  // the cold part of the exchange runs on a second stream, overlapped with the next sweep's panel gather
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_cold = nullptr;
  bool cold_pending = false;                 // ev_cold has been recorded and not yet waited for by a main sweep

  void* peer_xt[2][LUXB_MAX_PARTS]{};
  bool replica_stale = false;                // natural-order replica holds only this rank's slice (gathered on demand)
  // push apps
  uint32_t* d_cur = nullptr;       // [n_part] working labels of this partition
  uint64_t* d_out_end = nullptr;   // [nv] CSR-by-source end offsets over this partition's edges
  uint32_t* d_out_dst = nullptr;   // [e_part]
  void* d_big_list = nullptr;      // segments of hub sources' out-edge lists (push_big_kernel)
  uint32_t big_capacity = 0;
  unsigned char* d_fq_all = nullptr;  // every partition's frontier slot as exchanged
  unsigned char* d_fq_new = nullptr;  // this partition's slot under construction
  unsigned char* d_fq_tmp = nullptr;
  uint32_t* d_hdr_all = nullptr;      // [2 * P] gathered headers
  luxb::FrontierCtl* d_fctl = nullptr;  // device-side frontier finalisation flags (push.cuh)
  uint64_t* d_slot_off = nullptr;
  uint32_t* h_hdr = nullptr;          // pinned [2 * P]: type, count of the current frontier of every partition
  uint32_t* h_scratch = nullptr;      // pinned scratch (header readback)
  unsigned long long* d_counters = nullptr;  // [0] edges scanned by push kernels, [1] check mistakes
  // col_filter
  uint32_t* d_chunk_first = nullptr;
  uint32_t* d_chunk_vtx = nullptr;
  uint32_t n_chunks = 0;
  float* d_partial = nullptr;

  // communication
  luxb::ncclComm_t comm = nullptr;
  bool p2p_ready = false;
  void* peer_val[2][LUXB_MAX_PARTS]{};  // imported replicas of the peers (P2P exchange)
  void* peer_fq[LUXB_MAX_PARTS]{};      // imported frontier slot tables of the peers (CC / SSSP P2P push)
  uint32_t* d_sync = nullptr;
  // iteration barrier of the P2P paths without a library call: every rank owns P flag words, peers store their barrier
  // epoch into "their" word over NVLink and spin on their own (flag_barrier_kernel, build.cuh)
  uint32_t* d_flags = nullptr;
  void* peer_flags[LUXB_MAX_PARTS]{};
  uint32_t barrier_epoch = 0;
  uint32_t* h_barrier_err = nullptr;  // mapped pinned word: set when a peer did not arrive in time
  uint64_t barrier_timeout_ns = 30000000000ull;
  bool flag_barrier = true;           // LUXB_BARRIER=nccl: the 4-byte all-reduce of the communicator instead
  bool flag_barrier_all = false;      // LUXB_BARRIER=flag: also for the CC / SSSP / col_filter barriers
  bool direct_push = false;           // LUXB_PUSH=direct: owners store into EVERY rank's transfer array, no chunk pulls
  uint64_t ag_chunk = 0, hot_chunk = 0;  // equal chunk sizes (elements) of the balanced all-gather

  // source-blocked PageRank sweep (panel.cuh): hub destinations x hot source blocks in shared memory
  bool empties_done[2] = {false, false};  // value buffer k already holds update(identity) at the edge-less vertices
  bool seg_on = false;             // PageRank sweeps the flagged stream(s) of seg.cuh (sb_main [+ sb_panel])
  int seg_main_shape = 0, seg_panel_shape = 0;
  bool sb_on = false;
  PullLayout sb_main, sb_panel;
  PullLayout base_view;            // the canonical CSC seen as a layout by the merge-path sweep (pull.cuh)
  uint32_t sb_n_hub = 0, sb_n_blocks = 0, sb_bs = 0, sb_n_src = 0;
  int sb_shape = 0;
  uint32_t* d_hub_vtx = nullptr;
  uint32_t* d_hub_bits = nullptr;
  uint32_t* d_sb_partial = nullptr;  // [NV] raw panel reductions (4-byte Acc of the app's program)
  luxb::PanelBases sb_pb{};
  uint32_t sb_super_end[luxb::kPanelMaxBlocks]{};

  // launch configuration resolved once at open time (no getenv / function-static state on the hot path)
  int pull_ctas = 3;
  bool overlap_exchange = true;  // cold half of the PageRank exchange on the second stream (LUXB_OVERLAP=0: serialised)
  bool fused_fixup = true;  // one chained-scan launch instead of the three fix-up kernels (LUXB_FUSED_FIXUP=0: three)
  int panel_reserve_sms = 12;  // SMs the panel kernel leaves to the overlapped collective on several ranks
  int l2_hints = 1;   // LUXB_L2_HINTS: per-gather L2 eviction policies in the L1 sweep (hot evict_last, cold evict_first)
  PhaseTimer pt;

  // optional per-launch timing of the dominant kernel
  bool kernel_timing = false;
  std::vector<cudaEvent_t> kt_events;  // pairs
  size_t kt_used = 0;

  std::vector<void*> host_allocs;  // edge arrays living in mapped pinned host memory (cfg.zero_copy_edges)

  // stats / trace
  luxb_stats_t stats{};
  std::vector<uint64_t> trace_active;
  std::vector<int32_t> trace_pull;
};
