"""In-tree build of libksg_b200.so (nvcc, sm_100a only).

The library is the C ABI of ``include/ksg_b200.h``; it is built next to the sources
(``syntropy_b200/lib/``) so that it travels with a repo snapshot to the GPU box and
shows up as an in-tree native module.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libksg_b200.so")

OBJDIR = os.path.join(PKG, "build")
COMPILE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # exactness first: never contract a*b+c; the hot loops are add / compare only
    "-fmad=false",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [os.path.join(ROOT, "include", "ksg_b200.h")]
    for base, _, files in os.walk(CSRC):
        deps += [os.path.join(base, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    return deps


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libksg_b200.so cannot be built")


def _compile_one(args):
    src, obj, verbose = args
    cmd = [nvcc_path(), *COMPILE_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-I", os.path.join(ROOT, "include"), "-c", "-o", obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return proc.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources; return its path.
    One object per translation unit, compiled in parallel, then linked."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    jobs = []
    objs = []
    newest_header = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))
    for src in _sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((src, obj, verbose))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        logs = list(pool.map(_compile_one, jobs))
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc_path(), *LINK_FLAGS, "-o", LIB, *objs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
