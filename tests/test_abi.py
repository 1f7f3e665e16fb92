"""No-GPU checks of the drop-in boundary: libluxb.so builds/loads, exports every symbol include/lux_b200.h declares,
the host-side partitioner (pure host code behind the C ABI) matches the oracle, and the product refuses to run
without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lux_b200 as L
import oracle as O
from graphs import rmat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = L.load_library()
    names = L.declared_symbols()
    assert len(names) >= 24
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.luxb_version()


def test_header_cites_the_reference_interface():
    text = open(os.path.join(ROOT, "include", "lux_b200.h")).read()
    for cite in ("pull_model.inl", "push_model.inl", "pagerank_gpu.cu", "components_gpu.cu", "core/graph.h"):
        assert cite in text
    assert "torch" not in text.lower()  # plain pointers and sizes only


def test_product_sources_never_touch_the_oracle():
    """The product must not import, include, link or dlopen anything under oracle/ (comments may mention it)."""
    bad = re.compile(r"^\s*(import\s+oracle|from\s+oracle|#\s*include\s*[\"<][^\">]*oracle)|liblux_oracle|lo_[a-z_]+\(")
    for d in ("lux_b200", "lux_b200/csrc", "include"):
        for f in os.listdir(os.path.join(ROOT, d)):
            p = os.path.join(ROOT, d, f)
            if os.path.isfile(p) and f.endswith((".py", ".cu", ".cuh", ".h")):
                for ln, line in enumerate(open(p, errors="ignore"), 1):
                    if line.lstrip().startswith(("//", "*", "/*", "#  ", '"""')):
                        continue
                    assert not bad.search(line), "%s:%d %s" % (p, ln, line.strip())


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_host_partitioner_matches_oracle(P):
    row_end, src = rmat(12)
    cnt, rl, rr, cl = L.partition_csc(row_end, len(src), P)
    ocnt, orl, orr, ocl, _, _ = O.partition(row_end, len(src), P)
    assert cnt == ocnt
    assert np.array_equal(rl[:cnt], orl[:cnt]) and np.array_equal(rr[:cnt], orr[:cnt]) and np.array_equal(cl[:cnt], ocl[:cnt])


def test_partitioner_keeps_trailing_zero_indegree_vertices():
    # reference would assert (count 1 != P 2); we keep vertex 3 in an edge-free last partition
    row_end = np.array([2, 3, 5, 5], np.uint64)
    cnt, rl, rr, cl = L.partition_csc(row_end, 5, 2)
    assert cnt == 1 and (rl[0], rr[0]) == (0, 2) and (rl[1], rr[1]) == (3, 3) and cl[1] == 5


def test_bad_arguments_return_errors_not_exits():
    lib = L.load_library()
    with pytest.raises(L.LuxError):
        L.partition_csc(np.array([3, 2], np.uint64), 2, 1)  # decreasing row_end (pull_model.inl:100-101)
    assert lib.luxb_graph_info(None, None, None, None) < 0
    assert b"NULL" in lib.luxb_last_error()


def test_no_cpu_fallback_without_gpu(gpu_count):
    if gpu_count:
        pytest.skip("GPU present")
    with pytest.raises(L.LuxError, match="no CPU fallback"):
        L.pagerank(np.array([1, 2], np.uint64), np.array([1, 0], np.uint32), num_iter=1)
