"""In-tree build of ``libartist_b200.so`` (hand-written sm_100a CUDA + the C ABI of
``include/artist_b200.h``).  nvcc cross-compiles without a GPU; the ``.so`` is git-ignored but
travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libartist_b200.so")
# tuning only: AB200_LIB points at a pre-built variant of the library (tools/build_variants.sh)
LIB_OVERRIDE = os.environ.get("AB200_LIB")
SOURCES = ["trace.cu", "trace_bwd.cu", "nurbs.cu", "kinematics.cu", "blocking.cu", "geometry.cu", "flux.cu", "sampling.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or install the CUDA toolkit)")


def _source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/artist_b200.h"]
    for name in files:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path) and path.endswith((".cu", ".cuh", ".h")):
            with open(path, "rb") as fh:
                h.update(name.encode())
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("AB200_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if sources changed; return its path."""
    if LIB_OVERRIDE:
        if not os.path.exists(LIB_OVERRIDE):
            raise RuntimeError(f"AB200_LIB={LIB_OVERRIDE} does not exist")
        return LIB_OVERRIDE
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.stamp")
    digest = _source_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == digest:
                return LIB_PATH
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + os.environ.get("AB200_NVCC_EXTRA", "").split()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{out}")
        if verbose and out:
            print(out)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed: {' '.join(link)}\n{r.stdout}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
