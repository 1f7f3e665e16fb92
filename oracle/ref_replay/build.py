"""Compile the reference's own task bodies + kernels (from /root/reference, unmodified, nothing copied) behind the
Legion shim into oracle/_ref/libref_{pagerank,components,sssp}.so.  Needs /root/reference; the GPU box only uses the
prebuilt .so files (oracle/_ref/ travels with the snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "_ref")
APPS = {"pagerank": ("pagerank", "pagerank_gpu.cu", "REF_APP_PAGERANK"),
        "components": ("components", "components_gpu.cu", "REF_APP_COMPONENTS"),
        "sssp": ("sssp", "sssp_gpu.cu", "REF_APP_SSSP")}


def build():
    if not os.path.isdir(REF):
        return []
    os.makedirs(OUT, exist_ok=True)
    built = []
    for name, (d, cu, macro) in APPS.items():
        src = os.path.join(REF, d, cu)
        so = os.path.join(OUT, "libref_%s.so" % name)
        deps = [src, os.path.join(HERE, "replay.cu"), os.path.join(HERE, "shim", "legion.h"),
                os.path.join(HERE, "shim", "realm", "runtime_impl.h")]
        if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(p) for p in deps):
            built.append(so)
            continue
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-w", "-Xcompiler", "-fPIC", "-shared",
               "-D%s" % macro, '-DREF_CU="%s"' % src, "-I" + os.path.join(HERE, "shim"), "-I" + os.path.join(REF, d),
               "-o", so, os.path.join(HERE, "replay.cu")]
        subprocess.check_call(cmd)
        built.append(so)
    return built


if __name__ == "__main__":
    print("ref_replay:", build())
