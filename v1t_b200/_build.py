"""Build libv1t_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C-ABI)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libv1t_b200.so")
DIAG_LIB = os.path.join(PKG, "libv1t_b200_diag.so")  # micro-benchmarks / self-tests (csrc/diag), not the product ABI
STAMP = os.path.join(PKG, ".libv1t_b200.stamp")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _diag_sources():
    d = os.path.join(CSRC, "diag")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu")) if os.path.isdir(d) else []


def _digest():
    h = hashlib.sha256()
    files = _sources() + _diag_sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                                    if f.endswith((".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(PKG), "include", "v1t_b200.h"))
    files.append(os.path.join(os.path.dirname(PKG), "include", "v1t_b200_diag.h"))
    for f in files:
        h.update(os.path.basename(f).encode())  # names, not absolute paths: the tree moves (gpurun snapshot)
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    """True when the in-tree library was built from exactly the sources (and flags) that are in the tree now."""
    return (os.path.exists(LIB) and os.path.exists(DIAG_LIB) and os.path.exists(STAMP)
            and open(STAMP).read().strip() == _digest())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ and link the shared library.  Returns its path.  Serialised across processes
    (torchrun ranks) by a file lock; the ranks that waited find the library current and return."""
    import fcntl

    if not force and is_current():
        return LIB
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    with open(os.path.join(PKG, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    digest = _digest()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libv1t_b200.so")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in _sources() + _diag_sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, diag_objs = [], []
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        (diag_objs if os.path.dirname(src).endswith("diag") else objs).append(obj)
    for lib, group in ((LIB, objs), (DIAG_LIB, diag_objs)):
        cmd = [nvcc, "-shared", "-o", lib, *group, "-lcudart"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(objdir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(digest)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
