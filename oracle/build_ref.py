"""Builds what can be built of the UNMODIFIED reference, from the sources where they lie under /root/reference,
into oracle/_ref/ (git-ignored, travels to the GPU box).  Nothing is copied into the repo.

  converter   : tools/converter.cc (libc only) -> oracle/_ref/converter       [pins the .lux on-disk format]
  ref_replay  : the reference's own *_gpu.cu kernels behind a Legion shim      [see oracle/ref_replay/]

The rest of the reference cannot be built here: every other TU includes legion.h (core/graph.h:21) and the
legion/ submodule is empty (.SUBMODULES.json: status missing)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")


def build_converter():
    src = os.path.join(REF, "tools", "converter.cc")
    if not os.path.exists(src):
        return None
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "converter")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-w", "-o", exe, src])
    return exe


def main():
    built = [build_converter()]
    replay = os.path.join(HERE, "ref_replay", "build.py")
    if os.path.exists(replay):
        import runpy
        runpy.run_path(replay, run_name="__main__")
    print("oracle/_ref:", [b for b in built if b])


if __name__ == "__main__":
    sys.exit(main())
