"""Dev: hot-set coverage of RMAT-27 by out-degree + device L2 attributes (run on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lux_b200 as L
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, {k: getattr(p, k) for k in dir(p) if "persist" in k.lower() or "policy" in k.lower()})
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
for attr, name in ((108, "MaxPersistingL2CacheSize"), (109, "MaxAccessPolicyWindowSize")):
    v = ctypes.c_int(0); rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print(name, v.value)
scale = 27
with L.LuxGraph.from_rmat(scale, 1 << scale, 16 << scale, 27) as g:
    g.init()
    deg = g.out_degree()
d = np.sort(deg)[::-1].astype(np.int64)
cs = np.cumsum(d)
for mb in (16, 32, 64, 96, 128, 256):
    h = int(mb * 1e6 / 4)
    print("top %3d MB (%9d vertices, min degree %d): %.4f of all gathers" % (mb, h, d[h - 1], cs[h - 1] / cs[-1]))
print("vertices with degree>=1:", int((deg > 0).sum()), "deg>=2:", int((deg > 1).sum()))
