"""Dev tool: run the non-PageRank kernels once at config scale so that one `ncu --set full -k regex:...` pass can capture them:
C4 (SSSP, RMAT-24, start 0: push_relax / push_big / frontier_* / seg_tile_kernel<HopDist> pull sweeps) and C5 (col_filter,
NetFlix-scale: cf_chunk_kernel / cf_update_kernel).  Prints the timing lines; results are checked by tests/test_gpu_configs.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lux_b200 as L  # noqa: E402

with L.LuxGraph.from_rmat(24, 1 << 24, 16 << 24, 24, app=L.APP_SSSP, start=0) as g:
    g.init()
    it = g.run_to_convergence()
    st = g.stats()
    print(json.dumps(dict(config="C4 sssp rmat24", iters=it, total_ms=1e3 * st["loop_seconds"], pull_iterations=st["pull_iterations"],
                          mistakes=g.check())), flush=True)
with L.LuxGraph.from_bipartite(480189, 17770, 100480507, 5) as g:
    g.init()
    g.iterate(2)
    s0 = g.stats()
    g.iterate(3)
    s1 = g.stats()
    print(json.dumps(dict(config="C5 colfilter netflix-scale", ms_per_iter=1e3 * (s1["loop_seconds"] - s0["loop_seconds"]) / 3)), flush=True)
