"""The CLI drivers keep the reference's flags and stdout lines."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
from graphs import rmat

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "apps", "lux_cli.py")


def run(args):
    p = subprocess.run([sys.executable, CLI] + args, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_pagerank_cli(tmp_path):
    row_end, src = rmat(12)
    path, out = str(tmp_path / "g.lux"), str(tmp_path / "pr.npy")
    O.lux_write(path, row_end, src)
    txt = run(["pagerank", "-ll:gpu", "1", "-ll:fsize", "12000", "-ni", "7", "-file", path, "-out", out])
    assert re.search(r"\[Memory Setting\] Set ll:fsize >= \d+MB and ll:zsize >= \d+MB", txt)
    assert re.search(r"ELAPSED TIME = +\d+\.\d{7} s", txt)
    ref = O.pagerank(row_end, src, 7)
    assert np.allclose(np.load(out), ref, rtol=1e-6, atol=0)


def test_components_and_sssp_cli(tmp_path):
    row_end, src = rmat(12)
    path = str(tmp_path / "g.lux")
    O.lux_write(path, row_end, src)
    out = str(tmp_path / "cc.npy")
    txt = run(["components", "-ng", "1", "-file", path, "-check", "-out", out])
    assert "[PASS] Check task: rowLeft(0) numMistakes(0)" in txt
    assert np.array_equal(np.load(out), O.label_run(O.APP_CC, row_end, src)["labels"])
    out = str(tmp_path / "sssp.npy")
    txt = run(["sssp", "-ng", "1", "-file", path, "-start", "5", "-c", "-out", out])
    assert "[PASS] Check task" in txt
    assert np.array_equal(np.load(out), O.label_run(O.APP_SSSP, row_end, src, start=5)["labels"])


def test_colfilter_cli(tmp_path):
    row_end, src, w = O.gen_bipartite_csc(500, 40, 30000, 5)
    path, out = str(tmp_path / "r.lux"), str(tmp_path / "cf.npy")
    O.lux_write(path, row_end, src, w)
    run(["colfilter", "-ng", "1", "-ni", "3", "-file", path, "-out", out])
    assert np.allclose(np.load(out), O.colfilter(row_end, src, w, 3), rtol=2e-6, atol=0)
