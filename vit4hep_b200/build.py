"""Build libvit4hep_b200.so (the C-ABI library of include/vit4hep_b200.h) in-tree with nvcc.

sm_100a only: ``-gencode arch=compute_100a,code=sm_100a -lineinfo``.  nvcc cross-compiles
without a GPU, so this runs in the build container; the resulting .so travels to the GPU box
with the repo snapshot (it is git-ignored, not gpurun-ignored).

    python -m vit4hep_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(OUT_DIR, "libvit4hep_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    return h.hexdigest()


def _file_digest(path: str, salt: str) -> str:
    h = hashlib.sha256(salt.encode())
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _compile(nvcc: str, src: str, obj: str, digest: str, force: bool, verbose: bool) -> str:
    """Compile one translation unit unless its object is up to date; returns the ptxas log."""
    stamp, logf = obj + ".stamp", obj + ".log"
    if not force and all(os.path.isfile(p) for p in (obj, stamp, logf)):
        with open(stamp) as fh:
            if fh.read().strip() == digest:
                with open(logf) as lf:
                    return lf.read()
    cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
    with open(logf, "w") as fh:
        fh.write(r.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print(r.stderr)
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ (incrementally) and link the shared library."""
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = None
    objdir = os.path.join(OUT_DIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    srcs = sources()
    objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in srcs]
    salt = _headers_digest()
    digests = [_file_digest(s, salt) for s in srcs]
    total = hashlib.sha256("".join(digests).encode()).hexdigest()
    stamp = os.path.join(OUT_DIR, "build.stamp")
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == total:
                return LIB_PATH
    nvcc = _nvcc()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        logs = list(ex.map(lambda a: _compile(nvcc, a[0], a[1], a[2], force, verbose),
                           zip(srcs, objs, digests)))
    with open(os.path.join(OUT_DIR, "ptxas.log"), "w") as fh:
        fh.write("\n".join(logs))
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xlinker", "--no-undefined"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(total)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
