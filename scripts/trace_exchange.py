"""Exchange variants of the multi-GPU PageRank iteration side by side (run under torch.distributed.run, N >= 2):
the iteration barrier (flag kernel over NVLink vs the communicator's 4-byte all-reduce), the balanced two-step exchange vs
direct owner pushes, the overlap of the cold pull with the panel kernel, the SMs left to it.  Per variant: device and wall
time per iteration (max over ranks), per-rank phase times of one luxb_iterate(10) call, and a CRC of every rank's values
after the same number of iterations — the variants must agree bit for bit (the arithmetic is identical and deterministic).
usage: torchrun ... scripts/trace_exchange.py [scale] [variant,variant,...]   -> lines on stdout + gpurun_out/trace_exchange.txt"""
import json
import os
import sys
import time
import zlib

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lux_b200 as L  # noqa: E402

VARIANTS = {
    "flag": {},
    "direct": {"LUXB_PUSH": "direct"},
    "nccl": {"LUXB_BARRIER": "nccl"},
    "ov0": {"LUXB_OVERLAP": "0"},
    "res24": {"LUXB_PANEL_RESERVE_SMS": "24"},
    "direct_refsplit": {"LUXB_PUSH": "direct", "_balanced": "0"},
}
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 27
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["flag", "direct", "nccl", "ov0"]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
nv, ne, SEED = 1 << scale, 16 << scale, scale
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "trace_exchange.txt")


def allmax(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def emit(obj):
    if rank == 0:
        line = json.dumps(obj)
        print(line, flush=True)
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        with open(out_path, "a") as f:
            f.write(line + "\n")


ref_crc = None
for name in names:
    env = dict(VARIANTS[name])
    balanced = env.pop("_balanced", "1") == "1"
    for k in ("LUXB_PUSH", "LUXB_BARRIER", "LUXB_OVERLAP", "LUXB_PANEL_RESERVE_SMS", "LUXB_PHASE_TIMING"):
        os.environ.pop(k, None)
    os.environ.update(env)
    t_open = time.time()
    g = L.LuxGraph.from_rmat(scale, nv, ne, SEED, rank=rank, nranks=world, device=local, exchange=L.EXCHANGE_P2P, balanced=balanced)
    g.comm_init_torch()
    g.init()
    p2p = g.p2p_connect_torch()
    t_open = time.time() - t_open
    g.iterate(10)
    dist.barrier()
    s0 = g.stats()
    w0 = time.time()
    for _ in range(3):
        g.iterate(10)
    wall = (time.time() - w0) / 30
    s1 = g.stats()
    dev = allmax(s1["loop_seconds"] - s0["loop_seconds"]) / 30
    wall = allmax(wall)
    x = g.local_values()
    crc = zlib.crc32(np.ascontiguousarray(x).tobytes())
    crcs = [None] * world
    dist.all_gather_object(crcs, (crc, float(np.sum(x, dtype=np.float64))))
    if balanced and ref_crc is None:
        ref_crc = crcs
    same = (crcs == ref_crc) if balanced else None
    emit(dict(variant=name, env=env, balanced=balanced, n_gpus=world, scale=scale, p2p=bool(p2p), open_s=round(t_open, 1),
              device_ms_per_iter=1e3 * dev, wall_ms_per_iter=1e3 * wall, GTEPS=ne / dev / 1e9, same_values_as_first=same,
              value_sum=sum(c[1] for c in crcs), launches_per_iter=(s1["kernel_launches"] - s0["kernel_launches"]) / 30))
    os.environ["LUXB_PHASE_TIMING"] = "2"
    dist.barrier()
    g.iterate(10)  # one line per rank on stderr: the phase means of exactly these 10 iterations
    os.environ.pop("LUXB_PHASE_TIMING")
    g.close()
    dist.barrier()
