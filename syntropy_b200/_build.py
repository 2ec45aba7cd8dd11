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
    One object per translation unit, compiled in parallel, then linked.  Serialised across
    processes by a file lock (several torchrun ranks may find a stale library at once)."""
    if not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    import fcntl

    with open(os.path.join(LIBDIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # a sibling built it while we waited
                return LIB
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    from concurrent.futures import ThreadPoolExecutor

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
    # link under a temporary name and move it into place: another process (a sibling torchrun rank, a
    # snapshot of the tree) never sees a half-written library
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [nvcc_path(), *LINK_FLAGS, "-o", tmp, *objs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB)
    return LIB


# ------------------------------------------------------------------ the torch extension (host C++ only)
TORCH_EXT_SRC = os.path.join(PKG, "csrc_torch", "ksg_torch.cpp")
TORCH_EXT = os.path.join(LIBDIR, "ksg_torch_ext.so")


def torch_ext_is_stale() -> bool:
    if not os.path.exists(TORCH_EXT):
        return True
    t = os.path.getmtime(TORCH_EXT)
    return any(os.path.getmtime(d) > t for d in (TORCH_EXT_SRC, os.path.join(ROOT, "include", "ksg_b200.h")))


def build_torch_ext(force: bool = False) -> str:
    """Compile syntropy_b200/csrc_torch/ksg_torch.cpp into lib/ksg_torch_ext.so next to libksg_b200.so
    (g++ only: the extension holds no device code; it links the C ABI library through an $ORIGIN rpath)."""
    if not force and not torch_ext_is_stale():
        return TORCH_EXT
    build()  # the library it links against
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce

    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        raise RuntimeError("g++ not found; ksg_torch_ext.so cannot be built")
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = []
    for d in ce.include_paths(device_type="cuda") + [sysconfig.get_paths()["include"], os.path.join(ROOT, "include")]:
        inc += ["-isystem", d]
    abi = int(getattr(torch._C, "_GLIBCXX_USE_CXX11_ABI", True))
    tmp = TORCH_EXT + ".tmp%d" % os.getpid()
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=ksg_torch_ext",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, TORCH_EXT_SRC, "-o", tmp,
           "-L", LIBDIR, "-lksg_b200", "-L", torch_lib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda",
           "-ltorch_python", "-L", "/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib, "-Wl,--no-as-needed"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("torch extension build failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr[-4000:])
    os.replace(tmp, TORCH_EXT)
    return TORCH_EXT


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_ext(force="--force" in sys.argv))
