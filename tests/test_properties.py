"""Property tests (hypothesis, CPU): the product's host partitioner vs a line-by-line Python transcription of the
reference's greedy scan (pull_model.inl:108-131) on arbitrary degree sequences; oracle invariants on random graphs."""
import numpy as np
from hypothesis import given, settings, strategies as st

import lux_b200 as L
import oracle as O


def ref_partition(indeg, P):
    ne = int(sum(indeg))
    cap = (ne + P - 1) // P
    bounds, cnt, left = [], 0, 0
    for v, d in enumerate(indeg):
        cnt += d
        if cnt > cap:
            bounds.append((left, v))
            cnt, left = 0, v + 1
    if cnt > 0:
        bounds.append((left, len(indeg) - 1))
    return bounds


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=50), min_size=1, max_size=200), st.integers(min_value=1, max_value=8))
def test_partitioner_equals_reference_scan(indeg, P):
    if sum(indeg) == 0:
        indeg = list(indeg)
        indeg[0] = 1
    row_end = np.cumsum(np.array(indeg, np.uint64)).astype(np.uint64)
    ne = int(row_end[-1])
    want = ref_partition(indeg, P)
    cnt, rl, rr, cl = L.partition_csc(row_end, ne, P)
    assert cnt == len(want)  # what the reference would produce (it asserts cnt == P)
    for p, (a, b) in enumerate(want[:P]):
        if p < len(want) - 1 or len(want) == P or b == len(indeg) - 1:
            assert rl[p] == a
        assert cl[p] == (0 if a == 0 else int(row_end[a - 1]))
    # ours never drops a vertex: the non-empty partitions tile [0, nv) in order
    covered = []
    for p in range(P):
        n = (int(rr[p]) - int(rl[p]) + 1) & 0xFFFFFFFF
        if n and int(rl[p]) < len(indeg):
            covered.append((int(rl[p]), int(rr[p])))
    assert covered[0][0] == 0 and covered[-1][1] == len(indeg) - 1
    for (a0, b0), (a1, b1) in zip(covered, covered[1:]):
        assert a1 == b0 + 1
    if len(want) == P:  # the reference accepts the graph: bounds must be identical
        assert [(int(rl[p]), int(rr[p])) for p in range(P)] == want


@settings(max_examples=40, deadline=None)
@given(st.integers(min_value=2, max_value=60), st.integers(min_value=0, max_value=400), st.integers(min_value=0, max_value=2**31))
def test_oracle_invariants_on_random_graphs(nv, ne, seed):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, nv, ne).astype(np.uint32)
    d = rng.integers(0, nv, ne).astype(np.uint32)
    row_end, src = O.edges_to_csc(nv, s, d)
    # push == pull: the hybrid run and a pure Jacobi pull iteration reach the same fixed point
    r = O.label_run(O.APP_SSSP, row_end, src, start=0)
    lab = O.label_init(O.APP_SSSP, nv, 0)
    while True:
        new, changed = O.label_pull(O.APP_SSSP, row_end, src, lab)
        lab = new
        if changed == 0:
            break
    assert np.array_equal(r["labels"], lab)
    assert O.label_check(O.APP_SSSP, row_end, src, lab) == 0
    c = O.label_run(O.APP_CC, row_end, src)["labels"]
    assert O.label_check(O.APP_CC, row_end, src, c) == 0 and np.all(c >= np.arange(nv))
    # PageRank: stored value * max(deg,1) is the rank; every rank >= (1-alpha)/nv
    x = O.pagerank(row_end, src, 3)
    deg = O.out_degree(nv, src)
    assert np.all(x * np.maximum(deg, 1) >= np.float32(0.85) / np.float32(nv) * np.float32(0.999))
