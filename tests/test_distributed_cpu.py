"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: the NCCL-id bootstrap plumbing, agreement of the
partition table across ranks (product partitioner), and the per-iteration slice exchange protocol — exercised with
the oracle standing in for the device kernels (each rank computes only ITS destination range, then all-gathers)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lux_b200 as L
    import oracle as O
    try:
        # 1. id bootstrap: what comm_init_torch does before calling luxb_comm_init
        obj = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        assert obj[0] == bytes(range(128))
        # 2. every rank derives the same partition table from the same row_end (host code of libluxb)
        row_end, src = O.gen_rmat_csc(12, 4096, 65536, 27)
        nv, ne = len(row_end), len(src)
        cnt, rl, rr, cl = L.partition_csc(row_end, ne, world)
        tbl = torch.tensor(np.concatenate([rl, rr]).astype(np.int64))
        gathered = [torch.zeros_like(tbl) for _ in range(world)]
        dist.all_gather(gathered, tbl)
        assert all(torch.equal(g, tbl) for g in gathered) and cnt == world
        # 3. pull exchange protocol: compute own slice, all-gather unequal slices (pad to max), repeat
        deg = O.out_degree(nv, src)
        x = O.pagerank_init(deg)
        sizes = (rr.astype(np.int64) - rl.astype(np.int64) + 1)
        pad = int(sizes.max())
        for _ in range(4):
            mine = np.zeros(nv, np.float32)
            O.pagerank_iter(row_end, src, deg, x, int(rl[rank]), int(rr[rank]), out=mine)
            buf = torch.zeros(pad)
            buf[: sizes[rank]] = torch.from_numpy(mine[rl[rank]: rr[rank] + 1])
            outs = [torch.zeros(pad) for _ in range(world)]
            dist.all_gather(outs, buf)
            x = np.concatenate([outs[p][: sizes[p]].numpy() for p in range(world)])
        assert np.array_equal(x, O.pagerank(row_end, src, 4))
        # 4. push exchange protocol: per-partition label slices + global active count (halt test)
        lab = O.label_init(O.APP_SSSP, nv, 0)
        iters = 0
        while True:
            new, _ = O.label_pull(O.APP_SSSP, row_end, src, lab, int(rl[rank]), int(rr[rank]))
            changed = torch.tensor([int((new[rl[rank]: rr[rank] + 1] != lab[rl[rank]: rr[rank] + 1]).sum())])
            dist.all_reduce(changed)
            buf = torch.zeros(pad, dtype=torch.int64)
            buf[: sizes[rank]] = torch.from_numpy(new[rl[rank]: rr[rank] + 1].astype(np.int64))
            outs = [torch.zeros(pad, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(outs, buf)
            lab = np.concatenate([outs[p][: sizes[p]].numpy() for p in range(world)]).astype(np.uint32)
            iters += 1
            if int(changed) == 0:
                break
        ref = O.label_run(O.APP_SSSP, row_end, src, P=world, start=0)
        assert np.array_equal(lab, ref["labels"]) and iters == ref["iters"]
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_protocol():
    world, port = 2, 29641
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
