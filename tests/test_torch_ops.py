"""torch.ops.luxb.* (SURVEY §8 f4): registration is checked on CPU, results against the oracle on a GPU."""
import numpy as np
import pytest
import torch

import oracle as O
import lux_b200.torch_ops  # noqa: F401  (registers the ops)
from graphs import rmat


def test_ops_are_registered():
    for name in ("pagerank", "components", "sssp", "colfilter"):
        assert hasattr(torch.ops.luxb, name)
    assert "luxb::pagerank" in str(torch.ops.luxb.pagerank.default._schema)


@pytest.mark.gpu
def test_ops_match_the_oracle_on_gpu():
    row_end, src = rmat(12)
    re_t = torch.from_numpy(row_end.astype(np.int64)).cuda()
    src_t = torch.from_numpy(src.astype(np.int64)).cuda()
    pr = torch.ops.luxb.pagerank(re_t, src_t, 5)
    ref = O.pagerank(row_end, src, 5)
    assert pr.is_cuda and (np.abs(pr.cpu().numpy() - ref) / np.abs(ref)).max() <= 1e-6
    assert np.array_equal(torch.ops.luxb.sssp(re_t, src_t, 0).cpu().numpy(), O.label_run(O.APP_SSSP, row_end, src, start=0)["labels"])
    assert np.array_equal(torch.ops.luxb.components(re_t, src_t).cpu().numpy(), O.label_run(O.APP_CC, row_end, src)["labels"])
    re_b, src_b, w_b = O.gen_bipartite_csc(200, 30, 5000, 5)
    x = torch.ops.luxb.colfilter(torch.from_numpy(re_b.astype(np.int64)).cuda(), torch.from_numpy(src_b.astype(np.int64)).cuda(),
                                 torch.from_numpy(w_b).cuda(), 2)
    assert np.allclose(x.cpu().numpy(), O.colfilter(re_b, src_b, w_b, 2), rtol=2e-6, atol=0)
