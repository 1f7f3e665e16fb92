"""Worker for the multi-GPU parity test: run under torch.distributed.run, one rank per GPU.
Every rank opens ITS partition of the same deterministic graph, exchanges vertex values / frontiers each iteration
(NCCL all-gather or P2P stores), and rank 0 compares the result with the single-process CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lux_b200 as L  # noqa: E402
import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scale = int(os.environ.get("MGPU_SCALE", "16"))
    nv, ne, seed = 1 << scale, 16 << scale, 27
    row_end, src = O.gen_rmat_csc(scale, nv, ne, seed)
    ok = True

    def report(name, cond, extra=""):
        nonlocal ok
        if rank == 0:
            print("%s: %s %s" % (name, "OK" if cond else "FAIL", extra), flush=True)
        ok = ok and bool(cond)

    # ---- partition table must equal the reference greedy split (oracle) on every rank ----
    cnt, rl, rr, cl, _, _ = O.partition(row_end, ne, world)
    ref6 = O.pagerank(row_end, src, 6)
    for exchange, ename, sb in ((L.EXCHANGE_NCCL, "nccl", "0"), (L.EXCHANGE_P2P, "p2p", "0"), (L.EXCHANGE_P2P, "p2p+panel", "1"),
                                (L.EXCHANGE_NCCL, "nccl+panel", "1")):
        # sb = "1": force the source-blocked split on every rank (small blocks / thresholds), see tests/test_gpu_panel.py
        os.environ.update({"LUXB_SB": sb, "LUXB_SB_BS": "64", "LUXB_SB_MIN_INDEG": "4"})
        g = L.LuxGraph.from_rmat(scale, nv, ne, seed, rank=rank, nranks=world, device=local, exchange=exchange)
        b = g.bounds()
        report("partition[%s]" % ename, b["found"] == cnt and np.array_equal(b["row_left"], rl) and
               np.array_equal(b["row_right"], rr) and np.array_equal(b["col_left"], cl))
        g.comm_init_torch()
        g.init()
        if exchange != L.EXCHANGE_NCCL:
            report("p2p_connect[%s]" % ename, g.p2p_connect_torch())
        g.iterate(6)
        x = g.values()  # collective: completes the natural-order replica (PageRank exchanges only the packed values)
        err = (np.abs(x - ref6) / np.abs(ref6)).max()
        report("pagerank[%s] world=%d" % (ename, world), err <= 1e-6, "max rel err %.2e" % err)
        # every rank's own slice through the local API, then restart from it: 3 + 3 iterations = 6
        g2_lo, g2_n = g.local_range()
        mine = g.local_values()
        report("local_values[%s]" % ename, np.array_equal(mine, x[g2_lo:g2_lo + g2_n]))
        x3 = O.pagerank(row_end, src, 3)
        g.set_local_values(x3[g2_lo:g2_lo + g2_n])
        g.iterate(3)
        y = g.values()
        err = (np.abs(y - ref6) / np.abs(ref6)).max()
        report("set_local_values + 3 iterations[%s]" % ename, err <= 1e-6, "max rel err %.2e" % err)
        g.set_values(x3)
        g.iterate(3)
        err = (np.abs(g.values() - ref6) / np.abs(ref6)).max()
        report("set_values + 3 iterations[%s]" % ename, err <= 1e-6, "max rel err %.2e" % err)
        g.close()
        dist.barrier()
    os.environ["LUXB_SB"] = "0"

    # cost-balanced work split (cfg.balanced_split): other cut points, same answers; the reference's split is still reported
    for opener, oname in ((lambda: L.LuxGraph.from_rmat(scale, nv, ne, seed, rank=rank, nranks=world, device=local, exchange=L.EXCHANGE_P2P,
                                                       balanced=True), "rmat"),
                          (lambda: L.LuxGraph.from_csc(row_end, src, rank=rank, nranks=world, device=local, exchange=L.EXCHANGE_P2P,
                                                      balanced=True), "csc")):
        g = opener()
        b, w = g.bounds(), g.work_bounds()
        report("balanced[%s]: reference split still reported" % oname, np.array_equal(b["row_left"], rl) and np.array_equal(b["row_right"], rr))
        contiguous = w["row_left"][0] == 0 and all(int(w["row_right"][p]) + 1 == int(w["row_left"][p + 1]) for p in range(world - 1)) \
            and int(w["row_right"][world - 1]) == nv - 1
        report("balanced[%s]: contiguous cover, differs from the edge-balanced cuts" % oname,
               w["balanced"] and contiguous and not np.array_equal(w["row_left"], rl), str(w["row_left"]))
        g.comm_init_torch()
        g.init()
        g.p2p_connect_torch()
        g.iterate(6)
        err = (np.abs(g.values() - ref6) / np.abs(ref6)).max()
        report("pagerank[balanced %s] world=%d" % (oname, world), err <= 1e-6, "max rel err %.2e" % err)
        g.close()
        dist.barrier()

    # from host CSC arrays too (every rank passes the whole graph, keeps its slice)
    g = L.LuxGraph.from_csc(row_end, src, app=L.APP_PAGERANK, rank=rank, nranks=world, device=local)
    g.comm_init_torch()
    g.init()
    g.iterate(3)
    err = (np.abs(g.values() - O.pagerank(row_end, src, 3)) / np.abs(O.pagerank(row_end, src, 3))).max()
    report("pagerank[from_csc]", err <= 1e-6, "max rel err %.2e" % err)
    g.close()

    for app, oapp, name in ((L.APP_CC, O.APP_CC, "cc"), (L.APP_SSSP, O.APP_SSSP, "sssp")):
        ref = O.label_run(oapp, row_end, src, P=world, start=0)
        # frontier exchange by NCCL grouped broadcasts, then by direct P2P pushes into the peers' slot tables and label
        # replicas; the second pass also forces the source-blocked split in the pull sweeps
        for exchange, ename, sb in ((L.EXCHANGE_NCCL, "nccl", "0"), (L.EXCHANGE_P2P, "p2p push", "1")):
            os.environ.update({"LUXB_SB": sb, "LUXB_SB_BS": "64", "LUXB_SB_MIN_INDEG": "4"})
            g = L.LuxGraph.from_rmat(scale, nv, ne, seed, app=app, rank=rank, nranks=world, device=local, start=0, exchange=exchange)
            g.comm_init_torch()
            g.init()
            if exchange != L.EXCHANGE_NCCL:
                report("p2p_connect[%s %s]" % (name, ename), g.p2p_connect_torch())
            it = g.run_to_convergence()
            lab = g.values()
            bad = torch.tensor([g.check()], dtype=torch.int64, device="cuda")
            dist.all_reduce(bad)
            active, pull = g.trace()
            report("%s labels [%s] world=%d" % (name, ename, world), np.array_equal(lab, ref["labels"]) and int(bad) == 0)
            report("%s trace [%s]" % (name, ename), it == ref["iters"] and np.array_equal(active, ref["active"]) and np.array_equal(pull, ref["pull"]),
                   "iters %d vs %d" % (it, ref["iters"]))
            g.close()
            dist.barrier()
    os.environ["LUXB_SB"] = "0"

    # col_filter
    users, items, ratings = 4000, 200, 200000
    re_b, src_b, w_b = O.gen_bipartite_csc(users, items, ratings, 5)
    g = L.LuxGraph.from_bipartite(users, items, ratings, 5, rank=rank, nranks=world, device=local)
    g.comm_init_torch()
    g.init()
    g.iterate(3)
    x = g.values()
    ref = O.colfilter(re_b, src_b, w_b, 3)
    report("colfilter world=%d" % world, np.allclose(x, ref, rtol=2e-6, atol=0))
    g.close()

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_RESULT %s" % ("PASS" if int(flag) else "FAIL"), flush=True)
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
