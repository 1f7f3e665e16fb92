"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): partitions on separate ranks + per-iteration exchange
must reproduce the single-process oracle.  Launched exactly like the driver launches bench.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    return p.returncode, p.stdout + p.stderr


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world, gpu_count):
    if gpu_count < world:
        pytest.skip("needs %d GPUs, have %d" % (world, gpu_count))
    rc, out = _run(world, 29500 + world)
    assert rc == 0 and "MGPU_RESULT PASS" in out, out[-4000:]


def test_two_handles_two_threads_one_process(gpu_count):
    """The threading model INTEGRATION.md promises to a Lux maintainer: one host thread per GPU inside ONE process (Legion
    runs each partition's task body on that GPU's processor thread).  Two threads drive two handles concurrently —
    communicator set-up, init, iterations and the collective value read — and must reproduce the oracle."""
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs, have %d" % gpu_count)
    import threading
    import numpy as np
    sys.path.insert(0, ROOT)
    import lux_b200 as L
    import oracle as O
    scale = 15
    nv, ne = 1 << scale, 16 << scale
    row_end, src = O.gen_rmat_csc(scale, nv, ne, 27)
    uid = L.LuxGraph.comm_unique_id()
    out, err = [None, None], [None, None]

    def body(rank):
        try:
            g = L.LuxGraph.from_csc(row_end, src, app=L.APP_PAGERANK, rank=rank, nranks=2, device=rank, exchange=L.EXCHANGE_NCCL)
            g.comm_init(uid)
            g.init()
            g.iterate(5)
            out[rank] = g.values()
            labels = L.LuxGraph.from_csc(row_end, src, app=L.APP_SSSP, rank=rank, nranks=2, device=rank, start=0)
            labels.comm_init(uid2)
            labels.init()
            labels.run_to_convergence()
            out[rank] = (out[rank], labels.values())
            labels.close()
            g.close()
        except Exception as e:  # noqa: BLE001
            err[rank] = repr(e)

    uid2 = L.LuxGraph.comm_unique_id()
    threads = [threading.Thread(target=body, args=(r,)) for r in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert err == [None, None], err
    ref = O.pagerank(row_end, src, 5)
    ref_l = O.label_run(O.APP_SSSP, row_end, src, P=2, start=0)["labels"]
    for r in range(2):
        x, lab = out[r]
        assert (np.abs(x - ref) / np.abs(ref)).max() <= 1e-6
        assert np.array_equal(lab, ref_l)
