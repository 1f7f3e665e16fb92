"""PyTorch front-end of the C ABI: the four Lux apps as `torch.ops.luxb.*` custom ops taking the CSC as
torch tensors and returning torch tensors.  Plumbing only — every op opens a libluxb handle through the ctypes binding
(lux_b200/binding.py), runs the app on the CUDA device of the current torch context and copies the result back; no torch
kernel takes part in the computation, and there is no CPU fallback (the ops raise without a GPU).

    import lux_b200.torch_ops            # registers the ops
    ranks  = torch.ops.luxb.pagerank(row_end, src, 10)          # f32 [nv]  (rank / out-degree, pagerank_gpu.cu:98-100)
    labels = torch.ops.luxb.components(row_end, src)            # i64 [nv]  (max reaching id, components_gpu.cu:112-122)
    dist   = torch.ops.luxb.sssp(row_end, src, 0)               # i64 [nv]  (hop count, INF = nv, sssp_gpu.cu:122)
    x      = torch.ops.luxb.colfilter(row_end, src, weight, 10) # f32 [nv, 20]
row_end: int64 [nv] END offsets (the .lux convention); src: int64/int32 [ne]; weight: int32 [ne]."""
import numpy as np
import torch

from . import apps as _apps


def _np(t, dtype):
    return np.ascontiguousarray(t.detach().cpu().numpy()).astype(dtype, copy=False)


def _device_index(t):
    return t.device.index if t.is_cuda and t.device.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)


_lib = torch.library.Library("luxb", "DEF")
_lib.define("pagerank(Tensor row_end, Tensor src, int num_iter) -> Tensor")
_lib.define("components(Tensor row_end, Tensor src) -> Tensor")
_lib.define("sssp(Tensor row_end, Tensor src, int start) -> Tensor")
_lib.define("colfilter(Tensor row_end, Tensor src, Tensor weight, int num_iter) -> Tensor")


def _pagerank(row_end, src, num_iter):
    out = _apps.pagerank(_np(row_end, np.uint64), _np(src, np.uint32), num_iter=int(num_iter), device=_device_index(row_end))
    return torch.from_numpy(out).to(row_end.device)


def _components(row_end, src):
    out = _apps.components(_np(row_end, np.uint64), _np(src, np.uint32), device=_device_index(row_end))
    return torch.from_numpy(out["labels"].astype(np.int64)).to(row_end.device)


def _sssp(row_end, src, start):
    out = _apps.sssp(_np(row_end, np.uint64), _np(src, np.uint32), start=int(start), device=_device_index(row_end))
    return torch.from_numpy(out["labels"].astype(np.int64)).to(row_end.device)


def _colfilter(row_end, src, weight, num_iter):
    out = _apps.colfilter(_np(row_end, np.uint64), _np(src, np.uint32), _np(weight, np.int32), num_iter=int(num_iter),
                          device=_device_index(row_end))
    return torch.from_numpy(out).to(row_end.device)


for _name, _fn in (("pagerank", _pagerank), ("components", _components), ("sssp", _sssp), ("colfilter", _colfilter)):
    _lib.impl(_name, _fn, "CompositeExplicitAutograd")
