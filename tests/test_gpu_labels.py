"""GPU parity: CC (max label) and SSSP (= BFS depth) through the C ABI vs the CPU oracle — bit-exact labels,
identical per-iteration global active counts and pull/push direction decisions."""
import numpy as np
import pytest

import oracle as O
import lux_b200 as L
from graphs import ALL_SMALL, rmat, symmetrize

pytestmark = pytest.mark.gpu


def run_and_compare(app_fn, oracle_app, row_end, src, **kw):
    ref = O.label_run(oracle_app, row_end, src, P=1, **kw)
    out = app_fn(row_end, src, check=True, **kw)
    assert np.array_equal(out["labels"], ref["labels"])
    assert out["mistakes"] == 0
    assert out["iters"] == ref["iters"]
    active, pull = out["trace"]
    assert np.array_equal(active, ref["active"])
    assert np.array_equal(pull, ref["pull"])


@pytest.mark.parametrize("name", sorted(ALL_SMALL))
def test_cc_small_graphs(name):
    row_end, src = ALL_SMALL[name]()
    run_and_compare(L.components, O.APP_CC, row_end, src)


@pytest.mark.parametrize("name", sorted(ALL_SMALL))
def test_sssp_small_graphs(name):
    row_end, src = ALL_SMALL[name]()
    for start in (0, len(row_end) - 1):
        run_and_compare(L.sssp, O.APP_SSSP, row_end, src, start=start)


def test_cc_symmetric_rmat16():
    row_end, src = symmetrize(*rmat(16, ef=8))
    run_and_compare(L.components, O.APP_CC, row_end, src)


def test_sssp_rmat16_several_starts():
    row_end, src = rmat(16)
    for start in (0, 1, 12345):
        run_and_compare(L.sssp, O.APP_SSSP, row_end, src, start=start)


def test_device_generated_rmat18_cc_and_sssp():
    scale = 18
    nv, ne = 1 << scale, 16 << scale
    with L.LuxGraph.from_rmat(scale, nv, ne, 24, app=L.APP_SSSP, start=0) as g:
        row_end, src = g.local_csc()
        g.init()
        it = g.run_to_convergence()
        lab = g.values()
        assert g.check() == 0
    ref = O.label_run(O.APP_SSSP, row_end, src, P=1, start=0)
    assert it == ref["iters"] and np.array_equal(lab, ref["labels"])
    with L.LuxGraph.from_rmat(scale, nv, ne, 24, app=L.APP_CC) as g:
        g.init()
        it = g.run_to_convergence()
        lab = g.values()
        assert g.check() == 0
    ref = O.label_run(O.APP_CC, row_end, src, P=1)
    assert it == ref["iters"] and np.array_equal(lab, ref["labels"])


@pytest.mark.parametrize("app,oapp", [(L.APP_CC, O.APP_CC), (L.APP_SSSP, O.APP_SSSP)])
def test_check_counts_mistakes_like_the_reference_predicate(app, oapp):
    """luxb_check must COUNT violations of the reference's predicate (components_gpu.cu:786-790: label[dst] <
    label[src]; sssp_gpu.cu:792-796: label[src] != nv and label[dst] > label[src] + 1), not merely return 0 on a
    correct answer: corrupt k random labels of a converged result and compare the count with the oracle's."""
    row_end, src = rmat(14)
    nv = len(row_end)
    ref = O.label_run(oapp, row_end, src, P=1, start=0)["labels"]
    rng = np.random.default_rng(11)
    with L.LuxGraph.from_csc(row_end, src, app=app, start=0) as g:
        g.init()
        g.run_to_convergence()
        assert g.check() == 0
        for k in (1, 7, 300):
            bad = ref.copy()
            idx = rng.choice(nv, k, replace=False)
            if app == L.APP_CC:
                bad[idx] = rng.integers(0, nv, k).astype(np.uint32)       # arbitrary (mostly too small) labels
            else:
                bad[idx] = (bad[idx].astype(np.int64) + rng.integers(2, 9, k)).clip(0, nv).astype(np.uint32)  # too far
            want = O.label_check(oapp, row_end, src, bad)
            g.set_values(bad)
            got = g.check()
            assert got == want, (k, got, want)
            if k >= 7:
                assert want > 0
        g.set_values(ref)
        assert g.check() == 0
