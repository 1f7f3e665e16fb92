"""App drivers — the iteration loops of the reference's four top_level_tasks, on top of the C ABI.

pagerank   : pagerank/pagerank.cc:105-118     (-ni fixed iterations)
components : components/components.cc:108-135 (run until no partition reports an active vertex)
sssp       : sssp/sssp.cc                     (same loop, -start)
colfilter  : col_filter/colfilter.cc:71-81    (-ni fixed iterations)
Single-rank convenience wrappers; multi-GPU callers drive LuxGraph directly (see bench.py).
"""
from .binding import LuxGraph, APP_PAGERANK, APP_CC, APP_SSSP, APP_COLFILTER


def pagerank(row_end, src, num_iter=10, device=0):
    """Returns the array the reference holds in dist_lr[ni % 2]: rank / out-degree (pagerank_gpu.cu:98-100)."""
    with LuxGraph.from_csc(row_end, src, app=APP_PAGERANK, device=device) as g:
        g.init()
        g.iterate(num_iter)
        return g.values()


def components(row_end, src, device=0, check=False):
    with LuxGraph.from_csc(row_end, src, app=APP_CC, device=device) as g:
        g.init()
        iters = g.run_to_convergence()
        out = dict(labels=g.values(), iters=iters, trace=g.trace())
        if check:
            out["mistakes"] = g.check()
        return out


def sssp(row_end, src, start=0, device=0, check=False):
    with LuxGraph.from_csc(row_end, src, app=APP_SSSP, start=start, device=device) as g:
        g.init()
        iters = g.run_to_convergence()
        out = dict(labels=g.values(), iters=iters, trace=g.trace())
        if check:
            out["mistakes"] = g.check()
        return out


def colfilter(row_end, src, weight, num_iter=10, device=0):
    with LuxGraph.from_csc(row_end, src, weight, app=APP_COLFILTER, device=device) as g:
        g.init()
        g.iterate(num_iter)
        return g.values()
