"""Multi-GPU measurement of the BASELINE configs other than the headline (run under torch.distributed.run):
  C3 CC Twitter-scale (8 GPUs), C4 SSSP RMAT-24 (4 GPUs), C5 col_filter NetFlix-scale (8 GPUs).
usage: torchrun ... scripts/bench_configs_mgpu.py C3|C4|C5   -> one JSON line on rank 0 (max over ranks of device time)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lux_b200 as L  # noqa: E402

which = sys.argv[1]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def scale_of(nv):
    s = 1
    while (1 << s) < nv:
        s += 1
    return s


def maxreduce(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def sumreduce(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t)


if which in ("C3", "C4"):
    app, nv, ne, seed, name = ((L.APP_CC, 41652230, 1468365182, 3, "C3 components twitter-scale") if which == "C3"
                               else (L.APP_SSSP, 1 << 24, 16 << 24, 24, "C4 sssp rmat24 start 0"))
    exchange = L.EXCHANGE_NCCL if os.environ.get("LUXB_BENCH_EXCHANGE") == "nccl" else L.EXCHANGE_P2P
    runs = []
    for attempt in range(2):  # the second handle's run is the warm number (first use of a communicator / peer mapping is lazy)
        g = L.LuxGraph.from_rmat(scale_of(nv), nv, ne, seed, app=app, start=0, rank=rank, nranks=world, device=local, exchange=exchange)
        g.comm_init_torch()
        g.init()
        p2p = g.p2p_connect_torch() if exchange != L.EXCHANGE_NCCL else False
        dist.barrier()
        it = g.run_to_convergence()
        st = g.stats()
        bad = sumreduce(g.check())
        t = maxreduce(st["loop_seconds"])
        scanned = sumreduce(st["edges_processed"])
        active, pull = g.trace()
        runs.append(1e3 * t)
        g.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps(dict(config=name, n_gpus=world, nv=nv, ne=ne, iters=it, total_ms=runs[-1], first_run_ms=runs[0],
                              MTEPS_graph500=ne / (runs[-1] * 1e-3) / 1e6, MTEPS_edges_scanned=scanned / (runs[-1] * 1e-3) / 1e6,
                              pull_iterations=int(st["pull_iterations"]), check_mistakes=int(bad), active=[int(a) for a in active],
                              exchange="frontier P2P push" if p2p else "nccl grouped broadcasts")), flush=True)
else:
    users, items, ratings = 480189, 17770, 100480507
    g = L.LuxGraph.from_bipartite(users, items, ratings, 5, rank=rank, nranks=world, device=local, exchange=L.EXCHANGE_P2P)
    g.comm_init_torch()
    g.init()
    g.p2p_connect_torch()
    g.iterate(3)
    dist.barrier()
    s0 = g.stats()
    g.iterate(10)
    s1 = g.stats()
    t = maxreduce(s1["loop_seconds"] - s0["loop_seconds"])
    x = g.values()
    if rank == 0:
        ne = 2 * ratings
        print(json.dumps(dict(config="C5 colfilter netflix-scale", n_gpus=world, nv=users + items, ne=ne, iters=10,
                              ms_per_iter=1e3 * t / 10, MTEPS=ne * 10 / t / 1e6, finite=bool(np.isfinite(x).all()))), flush=True)
    g.close()
dist.barrier()
dist.destroy_process_group()
