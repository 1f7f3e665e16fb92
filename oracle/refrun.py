"""ctypes front-end of oracle/_ref/libref_*.so — the reference's OWN task bodies and CUDA kernels replayed behind a
Legion shim (oracle/ref_replay/).  TEST INFRASTRUCTURE: needs a GPU; used to (a) generate tests/golden/ref_replay_*.npz
on the GPU box (scripts/make_ref_golden.py) and (b) time the reference GPU path beside ours (bench extras)."""
import ctypes as C
import os

import numpy as np

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available(app="pagerank"):
    return os.path.exists(os.path.join(_REF, "libref_%s.so" % app))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def pagerank(row_end, src, ni):
    """Returns (values the reference holds in dist_lr[ni%2], wall ms of the ni PullAppTasks)."""
    lib = C.CDLL(os.path.join(_REF, "libref_pagerank.so"))
    row_end = np.ascontiguousarray(row_end, np.uint64)
    src = np.ascontiguousarray(src, np.uint32)
    out = np.zeros(len(row_end), np.float32)
    ms = C.c_double(0)
    rc = lib.ref_pagerank(C.c_uint32(len(row_end)), C.c_uint64(len(src)), _p(row_end), _p(src), C.c_int(ni), _p(out), C.byref(ms))
    assert rc == 0, rc
    return out, ms.value


def labels(app, row_end, src, start=0, max_iters=100000):
    """app in {'components','sssp'}.  Returns dict(labels, iters, active[iters], ms, mistakes)."""
    lib = C.CDLL(os.path.join(_REF, "libref_%s.so" % app))
    row_end = np.ascontiguousarray(row_end, np.uint64)
    src = np.ascontiguousarray(src, np.uint32)
    lab = np.zeros(len(row_end), np.uint32)
    active = np.zeros(max_iters, np.uint32)
    ms = C.c_double(0)
    bad = C.c_uint32(0)
    it = lib.ref_labels(C.c_uint32(len(row_end)), C.c_uint64(len(src)), _p(row_end), _p(src), C.c_uint32(start), _p(lab),
                        _p(active), C.c_int(max_iters), C.byref(ms), C.byref(bad))
    return dict(labels=lab, iters=it, active=active[:it].copy(), ms=ms.value, mistakes=bad.value)
