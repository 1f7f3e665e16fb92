"""ctypes/numpy front-end of oracle/lux_oracle.c (the CPU restatement of the reference semantics).

TEST INFRASTRUCTURE — see the header of lux_oracle.c.  Builds liblux_oracle.so on first use with the
committed Makefile (gcc is present on both the build container and the GPU box).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblux_oracle.so")
_lib = None

APP_CC, APP_SSSP = 1, 2
ALPHA = np.float32(0.15)
CF_K = 20
DENSE_BITMAP, SPARSE_QUEUE = 0x1234567, 0x7654321


def build(force=False):
    src = os.path.join(_HERE, "lux_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        L = _lib
        L.lo_splitmix64.restype = C.c_uint64
        L.lo_splitmix64.argtypes = [C.c_uint64]
        L.lo_edge_weight.restype = C.c_int32
        L.lo_edge_weight.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.lo_label_pull_range.restype = C.c_uint64
        L.lo_label_check.restype = C.c_uint64
        L.lo_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_threads():
    return lib().lo_num_threads()


def set_num_threads(n):
    lib().lo_set_num_threads(C.c_int(int(n)))


def splitmix64(x):
    return lib().lo_splitmix64(C.c_uint64(x & (2**64 - 1)))


def rmat_edge(seed, i, scale, nv):
    s, d = C.c_uint32(), C.c_uint32()
    lib().lo_rmat_edge(C.c_uint64(seed), C.c_uint64(i), C.c_int(scale), C.c_uint32(nv), C.byref(s), C.byref(d))
    return s.value, d.value


def gen_rmat_csc(scale, nv, ne, seed):
    """Canonical CSC ((dst,src)-sorted) of the deterministic RMAT graph. Returns (row_end u64[nv], src u32[ne])."""
    row_end = np.empty(nv, np.uint64)
    src = np.empty(max(ne, 1), np.uint32)[:ne]
    rc = lib().lo_gen_rmat_csc(C.c_int(scale), C.c_uint32(nv), C.c_uint64(ne), C.c_uint64(seed), _p(row_end), _p(src))
    assert rc == 0
    return row_end, src


def rmat_blocks(scale, nv, ne, seed, block_shift, block_sel, want_deg=True):
    """Sampled CSC of the RMAT graph: only destination blocks (2^block_shift ids) flagged in block_sel, compact
    numbering.  Returns dict(vid u32[n_local], row_end u64[n_local] (relative, inclusive), src u32[], deg u32[nv]|None).
    Two parallel scans of the edge counter; nothing of the product is involved."""
    block_sel = np.ascontiguousarray(block_sel, np.uint8)
    n_blocks = ((nv - 1) >> block_shift) + 1
    assert len(block_sel) == n_blocks
    blocks = np.nonzero(block_sel)[0].astype(np.int64)
    parts = [np.arange(b << block_shift, min((b + 1) << block_shift, nv), dtype=np.uint32) for b in blocks]
    vid = np.concatenate(parts) if parts else np.zeros(0, np.uint32)
    n_local = len(vid)
    deg = np.empty(nv, np.uint32) if want_deg else None
    indeg = np.empty(max(n_local, 1), np.uint32)[:n_local]
    rc = lib().lo_rmat_blocks_count(C.c_int(scale), C.c_uint32(nv), C.c_uint64(ne), C.c_uint64(seed), C.c_int(block_shift),
                                    _p(block_sel), _p(deg), _p(indeg), C.c_uint64(n_local))
    assert rc == 0, rc
    row_end = np.cumsum(indeg, dtype=np.uint64)
    ne_local = int(row_end[-1]) if n_local else 0
    src = np.empty(max(ne_local, 1), np.uint32)[:ne_local]
    rc = lib().lo_rmat_blocks_fill(C.c_int(scale), C.c_uint32(nv), C.c_uint64(ne), C.c_uint64(seed), C.c_int(block_shift),
                                   _p(block_sel), _p(row_end), C.c_uint64(n_local), _p(src))
    assert rc == 0, rc
    return dict(vid=vid, row_end=row_end, src=src, deg=deg)


def pagerank_iter_compact(nv, blk, deg, x_old, out=None):
    """One oracle PageRank iteration over the compact vertex set of rmat_blocks(); returns x_new[vid] (local order)."""
    n_local = len(blk["vid"])
    x_new = out if out is not None else np.empty(n_local, np.float32)
    lib().lo_pagerank_iter_compact(C.c_uint32(nv), C.c_uint64(n_local), _p(blk["row_end"]), _p(blk["src"]), _p(blk["vid"]),
                                   _p(deg), _p(x_old), _p(x_new))
    return x_new


def gen_bipartite_csc(users, items, ratings, seed):
    nv, ne = users + items, 2 * ratings
    row_end = np.empty(nv, np.uint64)
    src = np.empty(ne, np.uint32)
    w = np.empty(ne, np.int32)
    rc = lib().lo_gen_bipartite_csc(C.c_uint32(users), C.c_uint32(items), C.c_uint64(ratings), C.c_uint64(seed),
                                    _p(row_end), _p(src), _p(w))
    assert rc == 0
    return row_end, src, w


def edges_to_csc(nv, esrc, edst):
    esrc = np.ascontiguousarray(esrc, np.uint32)
    edst = np.ascontiguousarray(edst, np.uint32)
    ne = len(esrc)
    row_end = np.empty(nv, np.uint64)
    src = np.empty(max(ne, 1), np.uint32)[:ne]
    rc = lib().lo_edges_to_csc(C.c_uint32(nv), C.c_uint64(ne), _p(esrc), _p(edst), _p(row_end), _p(src))
    assert rc == 0, rc
    return row_end, src


def lux_write(path, row_end, src, weight=None):
    nv, ne = len(row_end), len(src)
    rc = lib().lo_lux_write(path.encode(), C.c_uint32(nv), C.c_uint64(ne), _p(row_end), _p(src), _p(weight))
    assert rc == 0


def lux_read(path, weighted=False):
    nv, ne = C.c_uint32(), C.c_uint64()
    rc = lib().lo_lux_read_header(path.encode(), C.byref(nv), C.byref(ne))
    assert rc == 0, rc
    row_end = np.empty(nv.value, np.uint64)
    src = np.empty(ne.value, np.uint32)
    w = np.empty(ne.value, np.int32) if weighted else None
    rc = lib().lo_lux_read(path.encode(), nv, ne, _p(row_end), _p(src), _p(w))
    assert rc == 0, rc
    return (row_end, src, w) if weighted else (row_end, src)


def partition(row_end, ne, P):
    """Reference greedy partitioner. Returns (count, row_left, row_right, col_left, fq_left, fq_right)."""
    nv = len(row_end)
    rl, rr = np.zeros(P, np.uint32), np.zeros(P, np.uint32)
    cl = np.zeros(P, np.uint64)
    fl, fr = np.zeros(P, np.uint64), np.zeros(P, np.uint64)
    cnt = lib().lo_partition(C.c_uint32(nv), C.c_uint64(ne), _p(row_end), C.c_int(P), _p(rl), _p(rr), _p(cl), _p(fl),
                             _p(fr))
    return cnt, rl, rr, cl, fl, fr


def out_degree(nv, src):
    deg = np.empty(nv, np.uint32)
    lib().lo_out_degree(C.c_uint32(nv), C.c_uint64(len(src)), _p(src), _p(deg))
    return deg


def pagerank_init(deg):
    x = np.empty(len(deg), np.float32)
    lib().lo_pagerank_init(C.c_uint32(len(deg)), _p(deg), _p(x))
    return x


def pagerank_iter(row_end, src, deg, x_old, v_lo=None, v_hi=None, out=None):
    nv = len(row_end)
    x_new = out if out is not None else np.zeros(nv, np.float32)
    if nv == 0:
        return x_new
    v_lo = 0 if v_lo is None else v_lo
    v_hi = nv - 1 if v_hi is None else v_hi
    lib().lo_pagerank_iter_range(C.c_uint32(nv), _p(row_end), _p(src), _p(deg), _p(x_old), _p(x_new),
                                 C.c_uint32(v_lo), C.c_uint32(v_hi))
    return x_new


def pagerank(row_end, src, iters):
    """Returns the array the reference would hold in dist_lr[ni%2]: rank / out-degree."""
    deg = out_degree(len(row_end), src)
    x = pagerank_init(deg)
    for _ in range(iters):
        x = pagerank_iter(row_end, src, deg, x)
    return x


def label_init(app, nv, start=0):
    lab = np.empty(nv, np.uint32)
    lib().lo_label_init(C.c_int(app), C.c_uint32(nv), C.c_uint32(start), _p(lab))
    return lab


def label_pull(app, row_end, src, old, v_lo=None, v_hi=None):
    nv = len(row_end)
    new = old.copy()
    v_lo = 0 if v_lo is None else v_lo
    v_hi = nv - 1 if v_hi is None else v_hi
    changed = lib().lo_label_pull_range(C.c_int(app), C.c_uint32(nv), _p(row_end), _p(src), _p(old), _p(new),
                                        C.c_uint32(v_lo), C.c_uint32(v_hi))
    return new, changed


def build_push_csr(row_end, src, v_lo, v_hi):
    """CSR-by-source over the edges of destination range [v_lo, v_hi]: (out_end u64[nv], out_dst u32[nedges])."""
    nv = len(row_end)
    e_lo = 0 if v_lo == 0 else int(row_end[v_lo - 1])
    e_hi = int(row_end[v_hi])
    out_end = np.empty(nv, np.uint64)
    out_dst = np.empty(max(e_hi - e_lo, 1), np.uint32)[: e_hi - e_lo]
    lib().lo_build_push_csr(C.c_uint32(nv), _p(row_end), _p(src), C.c_uint32(v_lo), C.c_uint32(v_hi), _p(out_end),
                            _p(out_dst))
    return out_end, out_dst


def label_run(app, row_end, src, P=1, start=0, max_iters=10000):
    """Run CC / SSSP with the reference's iteration structure.
    Returns dict(labels, iters, active[iters], pull[iters], ftype[iters,P])."""
    nv, ne = len(row_end), len(src)
    lab = np.empty(nv, np.uint32)
    active = np.zeros(max_iters, np.uint64)
    pull = np.zeros(max_iters, np.int32)
    ftype = np.zeros((max_iters, P), np.uint32)
    it = lib().lo_label_run(C.c_int(app), C.c_uint32(nv), C.c_uint64(ne), _p(row_end), _p(src), C.c_int(P),
                            C.c_uint32(start), _p(lab), C.c_int(max_iters), _p(active), _p(pull), _p(ftype))
    if it < 0:
        raise ValueError("reference partitioner does not yield P=%d partitions for this graph" % P)
    return dict(labels=lab, iters=it, active=active[:it].copy(), pull=pull[:it].copy(), ftype=ftype[:it].copy())


def label_check(app, row_end, src, label):
    return int(lib().lo_label_check(C.c_int(app), C.c_uint32(len(row_end)), _p(row_end), _p(src), _p(label)))


def cf_init(nv):
    x = np.empty((nv, CF_K), np.float32)
    lib().lo_cf_init(C.c_uint32(nv), _p(x))
    return x


def cf_iter(row_end, src, w, x_old, v_lo=None, v_hi=None):
    nv = len(row_end)
    x_new = np.zeros((nv, CF_K), np.float32)
    v_lo = 0 if v_lo is None else v_lo
    v_hi = nv - 1 if v_hi is None else v_hi
    lib().lo_cf_iter_range(C.c_uint32(nv), _p(row_end), _p(src), _p(w), _p(x_old), _p(x_new), C.c_uint32(v_lo),
                           C.c_uint32(v_hi))
    return x_new


def colfilter(row_end, src, w, iters):
    x = cf_init(len(row_end))
    for _ in range(iters):
        x = cf_iter(row_end, src, w, x)
    return x
