"""Run ON THE GPU BOX (gpurun): replays the reference's own kernels (oracle/_ref/libref_*.so, built here from
/root/reference by oracle/ref_replay/build.py) on small deterministic graphs, compares with the CPU oracle and with
libluxb, and writes the reference's outputs to gpurun_out/ref_replay_golden.npz.  The committed copy under
tests/golden/ pins the oracle by REFERENCE EXECUTION (tests/test_oracle.py::test_oracle_matches_reference_replay).
Also times the reference GPU path next to ours on a mid-size RMAT graph (gpurun_out/ref_replay_timing.json)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from oracle import refrun as R  # noqa: E402
import lux_b200 as L  # noqa: E402
from graphs import ALL_SMALL  # noqa: E402

out = {}
report = {}
names = ["hand5", "star", "rmat10", "rmat12_ragged_nv", "trailing_isolated", "two_components"]
for name in names:
    row_end, src = ALL_SMALL[name]()
    if len(src) == 0:
        continue
    pr_ref, _ = R.pagerank(row_end, src, 10)
    pr_or = O.pagerank(row_end, src, 10)
    rel = float((np.abs(pr_ref - pr_or) / np.abs(pr_or)).max())
    cc = R.labels("components", row_end, src)
    cc_or = O.label_run(O.APP_CC, row_end, src)
    ss = R.labels("sssp", row_end, src, start=0)
    ss_or = O.label_run(O.APP_SSSP, row_end, src, start=0)
    report[name] = dict(pagerank_max_rel_err_ref_vs_oracle=rel, cc_equal=bool(np.array_equal(cc["labels"], cc_or["labels"])),
                        cc_iters=(int(cc["iters"]), int(cc_or["iters"])), cc_mistakes=int(cc["mistakes"]),
                        sssp_equal=bool(np.array_equal(ss["labels"], ss_or["labels"])),
                        sssp_iters=(int(ss["iters"]), int(ss_or["iters"])), sssp_mistakes=int(ss["mistakes"]),
                        cc_active_equal=bool(np.array_equal(cc["active"], cc_or["active"])),
                        sssp_active_equal=bool(np.array_equal(ss["active"], ss_or["active"])))
    out[name + "_row_end"] = row_end
    out[name + "_src"] = src
    out[name + "_pagerank10"] = pr_ref
    out[name + "_cc"] = cc["labels"]
    out[name + "_cc_active"] = cc["active"]
    out[name + "_sssp0"] = ss["labels"]
    out[name + "_sssp0_active"] = ss["active"]
    print(name, report[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_replay_golden.npz"), **out)

# ---- timing: reference GPU path vs ours, same graph, same box ----
timing = {}
for scale in [int(s) for s in os.environ.get("REF_TIMING_SCALES", "22,24").split(",")]:
    nv, ne = 1 << scale, 16 << scale
    with L.LuxGraph.from_rmat(scale, nv, ne, 27) as g:
        row_end, src = g.local_csc()
        g.init()
        g.iterate(3)
        s0 = g.stats()
        g.iterate(10)
        s1 = g.stats()
        ours_ms = 1e3 * (s1["loop_seconds"] - s0["loop_seconds"]) / 10
        x_ours = g.values()
    t0 = time.time()
    x_ref, ref_ms = R.pagerank(row_end, src, 13)
    ref_total = time.time() - t0
    ccr = R.labels("components", row_end, src)
    with L.LuxGraph.from_csc(row_end, src, app=L.APP_CC) as g:
        g.init()
        it = g.run_to_convergence()
        cc_ms = 1e3 * g.stats()["loop_seconds"]
        cc_equal = bool(np.array_equal(g.values(), ccr["labels"]))
    timing["rmat%d" % scale] = dict(nv=nv, ne=ne, ours_pagerank_ms_per_iter=ours_ms, ref_pagerank_ms_per_iter=ref_ms / 13,
                                    ours_MTEPS=ne / ours_ms / 1e3, ref_MTEPS=ne / (ref_ms / 13) / 1e3,
                                    pagerank_max_rel_diff_ours_vs_ref=float((np.abs(x_ours - x_ref) / np.abs(x_ref)).max()),
                                    ref_init_plus_run_seconds=ref_total, ref_cc_ms=ccr["ms"], ref_cc_iters=int(ccr["iters"]),
                                    ours_cc_ms=cc_ms, ours_cc_iters=int(it), cc_labels_equal=cc_equal)
    print(scale, timing["rmat%d" % scale], flush=True)
json.dump(dict(report=report, timing=timing), open(os.path.join(ROOT, "gpurun_out", "ref_replay_timing.json"), "w"), indent=1)
