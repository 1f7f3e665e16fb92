"""In-tree build of libluxb.so (hand-written sm_100a CUDA + the C ABI).  nvcc cross-compiles without a GPU."""
import fcntl
import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libluxb.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + [
        os.path.join(_HERE, "..", "include", "lux_b200.h")]


HASH = LIB + ".srchash"


def source_hash():
    """Digest of everything the library is built from (source names + contents + compiler flags)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for s in _sources():
        h.update(os.path.basename(s).encode() + b"\0")
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale():
    """The library is current when the digest written next to it at build time equals the digest of the sources — file
    times do not survive every copy of the tree (a snapshot sent to a GPU box), contents do.  A library without a digest
    file (built by hand) falls back to the file-time rule."""
    if not os.path.exists(LIB):
        return True
    if os.path.exists(HASH):
        with open(HASH) as f:
            return f.read().strip() != source_hash()
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    """Compile lux_b200/csrc/api.cu (which includes every kernel header) into lux_b200/_lib/libluxb.so.
    Safe under torch.distributed.run: ranks serialise on a file lock, only the first one to get it compiles (to a
    temporary file that is renamed into place, so nobody ever dlopens a half-written library)."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another rank built it while we waited
                return LIB
            nvcc = os.environ.get("NVCC", "nvcc")
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC, "api.cu"), "-ldl"]
            try:
                digest = source_hash()  # of what the compiler is about to read
                subprocess.check_call(cmd, cwd=CSRC)
                os.replace(tmp, LIB)
                with open(HASH + ".tmp.%d" % os.getpid(), "w") as f:
                    f.write(digest + "\n")
                os.replace(HASH + ".tmp.%d" % os.getpid(), HASH)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
