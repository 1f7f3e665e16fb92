"""GPU parity: collaborative filtering (intended math) vs the CPU oracle.
Tolerance: f32 kernel vs an oracle that accumulates per-vertex in fp64 -> 2e-6 relative per value."""
import numpy as np
import pytest

import oracle as O
import lux_b200 as L

pytestmark = pytest.mark.gpu
REL_TOL = 2e-6


def test_colfilter_bipartite_small():
    row_end, src, w = O.gen_bipartite_csc(300, 40, 20000, 5)
    for ni in (1, 5):
        ref = O.colfilter(row_end, src, w, ni)
        gpu = L.colfilter(row_end, src, w, num_iter=ni)
        assert np.allclose(gpu, ref, rtol=REL_TOL, atol=0)


def test_colfilter_device_generator_and_hub_items():
    users, items, ratings = 20000, 300, 1500000  # items have in-degree >> chunk size
    row_end, src, w = O.gen_bipartite_csc(users, items, ratings, 5)
    with L.LuxGraph.from_bipartite(users, items, ratings, 5) as g:
        re_g, src_g, w_g = g.local_csc(weighted=True)
        assert np.array_equal(re_g, row_end) and np.array_equal(src_g, src) and np.array_equal(w_g, w)
        g.init()
        g.iterate(3)
        gpu = g.values()
    ref = O.colfilter(row_end, src, w, 3)
    assert np.allclose(gpu, ref, rtol=REL_TOL, atol=0)
