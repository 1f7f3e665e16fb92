"""GPU parity: PageRank pull path (libluxb through the C ABI) vs the CPU oracle.  Tolerance: 1e-6 relative
(BASELINE.json north_star) on the stored value rank/out-degree after every tested iteration count."""
import numpy as np
import pytest

import oracle as O
import lux_b200 as L
from graphs import ALL_SMALL, rmat

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6


def assert_close(gpu, ref):
    err = np.abs(gpu.astype(np.float64) - ref.astype(np.float64))
    bound = REL_TOL * np.abs(ref.astype(np.float64))
    bad = np.nonzero(err > bound)[0]
    assert bad.size == 0, "max rel err %.3e at %d (%d bad)" % ((err / np.maximum(np.abs(ref), 1e-300)).max(), bad[0], bad.size)


@pytest.mark.parametrize("name", sorted(ALL_SMALL))
def test_pagerank_small_graphs(name):
    row_end, src = ALL_SMALL[name]()
    for ni in (1, 2, 10):
        ref = O.pagerank(row_end, src, ni)
        gpu = L.pagerank(row_end, src, num_iter=ni)
        assert_close(gpu, ref)


def test_pagerank_rmat16_10iters():
    row_end, src = rmat(16)
    ref = O.pagerank(row_end, src, 10)
    gpu = L.pagerank(row_end, src, num_iter=10)
    assert_close(gpu, ref)


def test_pagerank_is_deterministic():
    row_end, src = rmat(14)
    a = L.pagerank(row_end, src, num_iter=5)
    b = L.pagerank(row_end, src, num_iter=5)
    assert np.array_equal(a, b)  # the reference's float atomicAdd is not (pagerank_gpu.cu:90)


def test_pagerank_device_rmat_generator_matches_oracle_generator():
    scale, nv = 13, 7000  # non power of two -> endpoint rejection path
    ne = 16 * nv
    row_end, src = O.gen_rmat_csc(scale, nv, ne, 3)
    with L.LuxGraph.from_rmat(scale, nv, ne, 3) as g:
        re_gpu, src_gpu = g.local_csc()
        assert np.array_equal(re_gpu, row_end)
        assert np.array_equal(src_gpu, src)
        g.init()
        g.iterate(10)
        assert_close(g.values(), O.pagerank(row_end, src, 10))


def test_pagerank_from_lux_file(tmp_path):
    row_end, src = rmat(11)
    path = str(tmp_path / "g.lux")
    O.lux_write(path, row_end, src)
    with L.LuxGraph.from_file(path) as g:
        g.init()
        g.iterate(3)
        assert_close(g.values(), O.pagerank(row_end, src, 3))


def test_pagerank_one_step_rmat22_many_fixup_blocks():
    """RMAT-22: 318 K warp tiles -> 1 242 fix-up blocks, i.e. more than the 1 024 threads of pull_fixup_blocks_kernel
    (serial chunks of 2 blocks per thread), and a hub whose in-edge list spans several fix-up blocks."""
    scale = 22
    nv, ne = 1 << scale, 16 << scale
    with L.LuxGraph.from_rmat(scale, nv, ne, 27) as g:
        row_end, src = g.local_csc()
        g.init()
        g.iterate(2)
        x2 = g.values()
        deg = g.out_degree()
        g.iterate(1)
        x3 = g.values()
    assert_close(x3, O.pagerank_iter(row_end, src, deg, x2))


def test_pagerank_one_step_property_rmat20():
    """Size-independent check usable at full scale: after k iterations on the device, one more device iteration
    must equal one oracle iteration applied to the device's own state; plus the linearity checksum
    sum_v s[v] == sum_u x[u] * outdeg[u]."""
    scale = 20
    nv, ne = 1 << scale, 16 << scale
    with L.LuxGraph.from_rmat(scale, nv, ne, 27) as g:
        row_end, src = g.local_csc()
        g.init()
        g.iterate(3)
        x3 = g.values()
        g.iterate(1)
        x4 = g.values()
    deg = O.out_degree(nv, src)
    assert_close(x4, O.pagerank_iter(row_end, src, deg, x3))
    init_rank = np.float32(0.85) / np.float32(nv)
    s_from_x4 = (x4.astype(np.float64) * np.maximum(deg, 1) - init_rank) / 0.15
    lhs = s_from_x4.sum()
    rhs = (x3.astype(np.float64) * deg).sum()
    assert abs(lhs - rhs) <= 1e-4 * abs(rhs)


def test_stats_and_errors():
    row_end, src = rmat(10)
    with L.LuxGraph.from_csc(row_end, src) as g:
        with pytest.raises(L.LuxError):
            g.iterate(1)  # before init
        g.init()
        g.iterate(4)
        st = g.stats()
        assert st["iterations"] == 4 and st["edges_processed"] == 4 * len(src) and st["kernel_launches"] >= 4
        with pytest.raises(L.LuxError):
            g.check()  # the reference has no PageRank check
    with pytest.raises(L.LuxError):
        L.LuxGraph.from_csc(np.array([3, 2], np.uint64), np.array([0, 1], np.uint32))  # decreasing row_end
    with pytest.raises(L.LuxError):
        L.LuxGraph.from_csc(np.array([1, 2], np.uint64), np.array([0, 9], np.uint32))  # src out of range


def test_zero_copy_edge_staging_matches():
    """cfg.zero_copy_edges: edge arrays stay in mapped pinned host memory (the -ll:zsize analogue) and are streamed
    over PCIe by the same TMA bulk copies; results must not change."""
    row_end, src = rmat(14)
    ref = O.pagerank(row_end, src, 4)
    with L.LuxGraph.from_csc(row_end, src, zero_copy=True) as g:
        g.init()
        g.iterate(4)
        assert_close(g.values(), ref)
    with L.LuxGraph.from_rmat(14, 1 << 14, 16 << 14, 27, app=L.APP_SSSP, zero_copy=True) as g:
        g.init()
        g.run_to_convergence()
        assert np.array_equal(g.values(), O.label_run(O.APP_SSSP, row_end, src, start=0)["labels"])


def test_balanced_work_split_follows_the_documented_cost_rule():
    """cfg.balanced_split (pull apps, nranks > 1): contiguous destination ranges cut where the running cost — 8 per
    vertex, 4 per edge into a hub (in-degree >= 64), 7 per other edge — passes k/P of the total; host path (from_csc) and
    device path (from_rmat) must agree with this restatement, and luxb_partition_bounds must keep reporting the
    reference's greedy split.  Opening a rank needs no communicator."""
    scale, nv = 15, 30000
    ne = 16 * nv
    row_end, src = O.gen_rmat_csc(scale, nv, ne, 27)
    indeg = np.diff(np.concatenate([[0], row_end]).astype(np.int64))
    cost = 8 + np.where(indeg >= 64, 4, 7) * indeg
    prefix = np.cumsum(cost)
    for P in (2, 5, 8):
        cuts, left = [], 0
        for p in range(P - 1):
            v = int(np.searchsorted(prefix * P, prefix[-1] * (p + 1), side="left"))
            v = max(v, left)
            cuts.append((left, v))
            left = v + 1
        cuts.append((left, nv - 1))
        cnt, rl, rr, cl, _, _ = O.partition(row_end, ne, P)
        for opener in (lambda r: L.LuxGraph.from_csc(row_end, src, rank=r, nranks=P, balanced=True),
                       lambda r: L.LuxGraph.from_rmat(scale, nv, ne, 27, rank=r, nranks=P, balanced=True)):
            with opener(P - 1) as g:
                w, b = g.work_bounds(), g.bounds()
                assert w["balanced"]
                assert [(int(a), int(c)) for a, c in zip(w["row_left"], w["row_right"])] == cuts
                assert np.array_equal(b["row_left"], rl) and np.array_equal(b["row_right"], rr) and np.array_equal(b["col_left"], cl)
                lo, n = g.local_range()
                assert (lo, lo + n - 1) == cuts[P - 1]
