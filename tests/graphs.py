"""Small deterministic input graphs shared by the oracle tests and the GPU parity tests (built with the oracle)."""
import numpy as np

import oracle as O


def hand5():
    """The 5-edge graph whose .lux bytes SURVEY §8f-1 verified against tools/converter.cc:
    edges {0->1, 1->2, 2->0, 3->0, 0->2}."""
    return O.edges_to_csc(4, [0, 1, 2, 3, 0], [1, 2, 0, 0, 2])


def star(n=10000, both=True):
    """Hub: every vertex -> 0 (in-degree n-1 spans several merge tiles); optionally 0 -> every vertex."""
    s = list(range(1, n))
    d = [0] * (n - 1)
    if both:
        s += [0] * (n - 1)
        d += list(range(1, n))
    return O.edges_to_csc(n, s, d)


def chain(n=3000, symmetric=False):
    s = list(range(0, n - 1))
    d = list(range(1, n))
    if symmetric:
        s, d = s + d, d + s
    return O.edges_to_csc(n, s, d)


def no_edges(n=7):
    return np.zeros(n, np.uint64), np.zeros(0, np.uint32)


def trailing_isolated(n_core=500, n_iso=5000, seed=9):
    """RMAT core followed by a long tail of vertices with no edges at all (ragged tiles: all-vertex tiles)."""
    re, src = O.gen_rmat_csc(9, n_core, 8 * n_core, seed)
    re2 = np.concatenate([re, np.full(n_iso, re[-1], np.uint64)])
    return re2, src


def rmat(scale, ef=16, seed=27, nv=None):
    nv = (1 << scale) if nv is None else nv
    return O.gen_rmat_csc(scale, nv, ef * nv, seed)


def symmetrize(row_end, src):
    nv = len(row_end)
    dst = np.repeat(np.arange(nv, dtype=np.uint32), np.diff(np.concatenate([[0], row_end]).astype(np.int64)))
    return O.edges_to_csc(nv, np.concatenate([src, dst]), np.concatenate([dst, src]))


def two_components(n=2000):
    """Two disjoint symmetric chains + isolated vertices: CC must label them max-id per component."""
    h = n // 2
    s = list(range(0, h - 1)) + list(range(h, n - 11))
    d = list(range(1, h)) + list(range(h + 1, n - 10))
    return O.edges_to_csc(n, s + d, d + s)


ALL_SMALL = {
    "hand5": hand5,
    "star": star,
    "chain": chain,
    "no_edges": no_edges,
    "trailing_isolated": trailing_isolated,
    "rmat10": lambda: rmat(10),
    "rmat12_ragged_nv": lambda: rmat(12, nv=3000),
    "two_components": two_components,
}
