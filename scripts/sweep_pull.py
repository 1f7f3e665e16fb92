"""Dev tool: sweep the merge-path tile shapes of the pull kernel (LUXB_PULL_SHAPE) on one GPU.
Correctness of every shape is checked against the oracle on a small graph first."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lux_b200 as L  # noqa: E402
import oracle as O  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 27
shapes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
hots = sys.argv[3].split(",") if len(sys.argv) > 3 else ["64"]
ctas = sys.argv[4].split(",") if len(sys.argv) > 4 else ["2"]
row_end, src = O.gen_rmat_csc(14, 1 << 14, 16 << 14, 27)
ref = O.pagerank(row_end, src, 5)
for sh, hot, ct in [(a, b, c) for a in shapes for b in hots for c in ctas]:
    os.environ["LUXB_PULL_SHAPE"] = str(sh)
    os.environ["LUXB_HOT_MB"] = hot
    os.environ["LUXB_PULL_CTAS"] = ct
    gpu = L.pagerank(row_end, src, num_iter=5)
    err = (np.abs(gpu - ref) / np.abs(ref)).max()
    nv, ne = 1 << scale, 16 << scale
    with L.LuxGraph.from_rmat(scale, nv, ne, 27) as g:
        g.init()
        g.iterate(5)
        gm0, gm1 = g.debug_gather_ms(False), g.debug_gather_ms(True)
        g.enable_kernel_timing(True)
        s0 = g.stats()
        g.iterate(20)
        s1 = g.stats()
    k = (s1["dominant_kernel_seconds"] - s0["dominant_kernel_seconds"]) / 20
    t = (s1["loop_seconds"] - s0["loop_seconds"]) / 20
    print("ctas %s hot %s MB shape %d: parity max rel err %.2e | scale %d kernel %.3f ms, iter %.3f ms, %.1f GTEPS, algo %.0f GB/s | bare gather natural %.2f ms packed %.2f ms" % (
        ct, hot, sh, err, scale, k * 1e3, t * 1e3, ne / t / 1e9, (8 * ne + 16 * nv) / k / 1e9, gm0, gm1), flush=True)
