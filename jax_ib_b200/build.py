"""Builds the native library IN-TREE with nvcc for sm_100a (no torch, no JIT cache).

    python -m jax_ib_b200.build [--force] [--verbose]

Output: jax_ib_b200/_native/libjaxib_b200.so (git-ignored; travels to the GPU box with the tree).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_native")
LIB = os.path.join(OUT_DIR, "libjaxib_b200.so")
STAMP = os.path.join(OUT_DIR, "build.stamp")
SOURCES = ["capi.cu", "stencil.cu", "stencil_fast.cu", "stencil_ring.cu", "ib.cu", "poisson.cu", "poisson_fast.cu", "poisson_team.cu", "poisson_team_rows.cu", "poisson_sweep.cu", "tracer.cu", "slab.cu"]
HEADERS = ["common.cuh", "fft.cuh", "poisson_fast.cuh", os.path.join("..", "..", "include", "jaxib_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-DJAXIB_ARCH=100",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (no CPU fallback exists)")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest() -> str:
    h = hashlib.sha256()
    for name in _sources() + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _fresh(digest: str) -> bool:
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    digest = _digest()
    if not force and _fresh(digest):
        return LIB
    import fcntl

    with open(os.path.join(OUT_DIR, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)  # one builder at a time (torchrun ranks share the tree)
        if not force and _fresh(digest):  # another rank built it while we waited
            return LIB
        return _build_locked(digest, verbose)


def _build_locked(digest: str, verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out, file=sys.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    if os.environ.get("JAXIB_WITH_NCCL", "1") == "1" and os.path.exists(os.path.join(CSRC, "slab.cu")):
        link += _nccl_link_flags()
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


def _nccl_link_flags():
    return []


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
