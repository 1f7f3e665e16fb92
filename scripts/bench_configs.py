"""Measures the other BASELINE.json configs on ONE GPU (they are parity-test cases, not the bench headline):
  C1 PageRank, Indochina-scale synthetic (7,414,866 V / 194,109,311 E), -ni 10
  C3 CC, Twitter-2010-scale synthetic (41,652,230 V / 1,468,365,182 E)        [BASELINE quotes 8 GPUs]
  C4 SSSP (hop count), RMAT-24 edge factor 16, -start 0                         [BASELINE quotes 4 GPUs]
  C5 col_filter, NetFlix-scale bipartite (480,189 + 17,770 V / 200,961,014 E)   [BASELINE quotes 8 GPUs]
Each line: config, MTEPS (ne * iterations / loop time for PR/CF; ne / total time for CC/SSSP, Graph500 style, plus
edges actually scanned), iterations, and the size-independent parity property that was checked at full size
(the reference's own -check invariant = 0 mistakes; PageRank linearity checksum)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lux_b200 as L  # noqa: E402

which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C1", "C3", "C4", "C5"]
out = []
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    PEAK = 6650.0  # B200_PROFILING.md fallback


def scale_of(nv):
    s = 1
    while (1 << s) < nv:
        s += 1
    return s


if "C1" in which:
    nv, ne = 7414866, 194109311
    with L.LuxGraph.from_rmat(scale_of(nv), nv, ne, 1) as g:
        g.init()
        g.iterate(10)
        s0 = g.stats()
        g.iterate(10)
        s1 = g.stats()
        x_prev = None
        t = s1["loop_seconds"] - s0["loop_seconds"]
        deg = g.out_degree()
        x = g.values()
        g.iterate(1)
        y = g.values()
    init = np.float32(0.85) / np.float32(nv)
    lhs = ((y.astype(np.float64) * np.maximum(deg, 1) - init) / 0.15).sum()
    rhs = (x.astype(np.float64) * deg).sum()
    out.append(dict(config="C1 pagerank indochina-scale", nv=nv, ne=ne, iters=10, ms_per_iter=1e3 * t / 10,
                    MTEPS=ne * 10 / t / 1e6, linearity_checksum_rel_err=abs(lhs - rhs) / abs(rhs),
                    roofline=dict(bound="hbm", algorithmic_bytes_per_iter=8 * ne + 16 * nv, achieved_GBps=(8 * ne + 16 * nv) * 10 / t / 1e9,
                                  peak_GBps=PEAK, frac=(8 * ne + 16 * nv) * 10 / t / 1e9 / PEAK, note="whole iteration, SURVEY 8d bytes")))
    print(json.dumps(out[-1]), flush=True)

for tag, app, nv, ne, seed, name in (("C3", L.APP_CC, 41652230, 1468365182, 3, "C3 components twitter-scale"),
                                     ("C4", L.APP_SSSP, 1 << 24, 16 << 24, 24, "C4 sssp rmat24 start 0")):
    if tag not in which:
        continue
    with L.LuxGraph.from_rmat(scale_of(nv), nv, ne, seed, app=app, start=0) as g:
        g.init()
        it = g.run_to_convergence()
        st = g.stats()
        bad = g.check()
        active, pull = g.trace()
        lab = g.values()
    t = st["loop_seconds"]
    reached = int((lab != nv).sum()) if app == L.APP_SSSP else int(np.unique(lab).size)
    out.append(dict(config=name, nv=nv, ne=ne, iters=it, total_ms=1e3 * t, MTEPS_graph500=ne / t / 1e6,
                    MTEPS_edges_scanned=st["edges_processed"] / t / 1e6, pull_iterations=st["pull_iterations"],
                    check_mistakes=bad, active=[int(a) for a in active], reached_or_components=reached))
    print(json.dumps(out[-1]), flush=True)

if "C5" in which:
    users, items, ratings = 480189, 17770, 100480507
    with L.LuxGraph.from_bipartite(users, items, ratings, 5) as g:
        g.init()
        g.iterate(3)
        s0 = g.stats()
        g.iterate(10)
        s1 = g.stats()
        x = g.values()
    t = s1["loop_seconds"] - s0["loop_seconds"]
    ne = 2 * ratings
    nvv = users + items
    out.append(dict(config="C5 colfilter netflix-scale", nv=nvv, ne=ne, iters=10, ms_per_iter=1e3 * t / 10,
                    MTEPS=ne * 10 / t / 1e6, finite=bool(np.isfinite(x).all()), mean_abs=float(np.abs(x).mean()),
                    roofline=dict(bound="hbm", algorithmic_bytes_per_iter=8 * ne + 168 * nvv, achieved_GBps=(8 * ne + 168 * nvv) * 10 / t / 1e9,
                                  peak_GBps=PEAK, frac=(8 * ne + 168 * nvv) * 10 / t / 1e9 / PEAK,
                                  l2_gather_bytes_per_iter=80 * ne, l2_gather_GBps=80 * ne * 10 / t / 1e9,
                                  note="secondary bound (SURVEY 8d): the 80-byte vector gathers are 3 L1 sector requests each = "
                                       "%.0f M sectors per iteration; at the measured 290 G sectors/s wall that is %.2f ms" % (
                                           3 * ne / 1e6, 3 * ne / 290e9 * 1e3))))
    print(json.dumps(out[-1]), flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs_1gpu.json", "w"), indent=1)
