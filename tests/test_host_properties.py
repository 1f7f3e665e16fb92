"""Property tests (hypothesis, CPU only) of the product's HOST entry points against the oracle — the pieces of the path
that run without a device: the partitioner (Graph::Graph, pull_model.inl:97-131 ≡ push_model.inl:367-423), the .lux
writer (tools/converter.cc:98-124) and the edge-list converter (tools/converter.cc:72-130).  Random ragged inputs: vertices
without in-edges, empty graphs, more partitions than non-empty vertices, hubs holding most of the edges."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import lux_b200 as L
import oracle as O

degrees = st.lists(st.one_of(st.just(0), st.integers(0, 6), st.integers(0, 400)), min_size=1, max_size=200)


@settings(max_examples=150, deadline=None)
@given(deg=degrees, P=st.integers(1, 16))
def test_partition_table_equals_the_oracle_scan(deg, P):
    row_end = np.cumsum(np.array(deg, np.uint64)).astype(np.uint64)
    ne = int(row_end[-1])
    cnt, rl, rr, cl = L.partition_csc(row_end, ne, P)
    ocnt, orl, orr, ocl, _, _ = O.partition(row_end, ne, P)
    assert cnt == ocnt
    assert np.array_equal(rl[:cnt], orl[:cnt]) and np.array_equal(rr[:cnt], orr[:cnt]) and np.array_equal(cl[:cnt], ocl[:cnt])
    # what every consumer relies on: the first cnt ranges are contiguous, start at 0, colLeft = edges before the range
    if cnt:
        assert rl[0] == 0 and all(int(rl[k + 1]) == int(rr[k]) + 1 for k in range(cnt - 1))
        assert all(int(cl[k]) == (int(row_end[int(rl[k]) - 1]) if rl[k] else 0) for k in range(cnt))


@settings(max_examples=60, deadline=None)
@given(nv=st.integers(1, 60), edges=st.lists(st.tuples(st.integers(0, 59), st.integers(0, 59)), min_size=0, max_size=300),
       weighted=st.booleans())
def test_lux_writer_bytes_equal_the_oracle_writer(tmp_path_factory, nv, edges, weighted):
    edges = [(s % nv, d % nv) for s, d in edges]
    s = np.array([e[0] for e in edges], np.uint32)
    d = np.array([e[1] for e in edges], np.uint32)
    row_end, src = O.edges_to_csc(nv, s, d)
    w = (np.arange(len(src), dtype=np.int32) % 5 + 1) if weighted else None
    tmp = tmp_path_factory.mktemp("lux")
    a, b = str(tmp / "a.lux"), str(tmp / "b.lux")
    L.write_lux(a, row_end, src, w)
    O.lux_write(b, row_end, src, w)
    assert open(a, "rb").read() == open(b, "rb").read()


@settings(max_examples=40, deadline=None)
@given(nv=st.integers(1, 80), edges=st.lists(st.tuples(st.integers(0, 79), st.integers(0, 79)), min_size=1, max_size=400))
def test_converter_output_is_the_canonical_lux_of_the_edge_list(tmp_path_factory, nv, edges):
    edges = [(s % nv, d % nv) for s, d in edges]
    tmp = tmp_path_factory.mktemp("conv")
    txt = tmp / "edges.txt"
    txt.write_text("".join("%d %d\n" % e for e in edges))
    out = str(tmp / "out.lux")
    L.convert_edgelist(str(txt), out, nv, len(edges))
    row_end, src = O.edges_to_csc(nv, np.array([e[0] for e in edges], np.uint32), np.array([e[1] for e in edges], np.uint32))
    canon = str(tmp / "canon.lux")
    O.lux_write(canon, row_end, src)
    assert open(out, "rb").read() == open(canon, "rb").read()
    # and it loads back as the same CSC (the oracle's reader = the reference's load path, pull_model.inl:253-320)
    re2, src2 = O.lux_read(out)
    assert np.array_equal(re2, row_end) and np.array_equal(src2, src)


def test_partitioner_rejects_what_the_reference_asserts_on():
    with pytest.raises(L.LuxError):
        L.partition_csc(np.array([5, 4, 9], np.uint64), 9, 2)     # decreasing offsets (pull_model.inl:100-101)
    with pytest.raises(L.LuxError):
        L.partition_csc(np.array([1, 2, 3], np.uint64), 3, 0)     # no partitions
