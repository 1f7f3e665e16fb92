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
