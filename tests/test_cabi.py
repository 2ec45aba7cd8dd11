"""CPU-side checks of the boundary: the library builds for sm_100a, loads, and exports
every symbol include/ksg_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ksg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ksg_[a-z0-9_]+)\s*\(", text)))


def test_header_lists_the_expected_entry_points():
    names = declared_symbols()
    for must in ("ksg_knn_radius", "ksg_count_subspaces", "ksg_psi_epilogue", "ksg_paircount_fixed_radius",
                 "ksg_psi_table", "ksg_sum_f64", "ksg_last_error"):
        assert must in names


def test_library_builds_loads_and_exports_every_symbol():
    from syntropy_b200 import _build, _cabi

    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in ksg_b200.h but not exported"
    typed = _cabi.load()
    assert set(_cabi.SIGNATURES) == set(declared_symbols())
    assert typed.ksg_version() >= 1000
    # workspace sizing is host-only arithmetic and must work without a device
    assert typed.ksg_knn_radius_workspace(1_000_000, 125_000, 2, 6) > 0
    assert typed.ksg_sum_workspace(10) > 0


def test_bad_arguments_return_codes_not_crashes():
    from syntropy_b200 import _cabi

    lib = _cabi.load()
    rc = lib.ksg_knn_radius(None, 0, 0, 2, None, 0, 0, 0, 3, None, None, None, 0, None)
    assert rc == -1 and b"ksg_knn_radius" in lib.ksg_last_error()
    rc = lib.ksg_psi_table(None, 10, None)
    assert rc == -1


def test_no_cpu_fallback_without_cuda():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from syntropy_b200 import knn

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        knn.mutual_information((0,), (1,), 3, np.zeros((2, 10)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "syntropy_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|ksg_oracle|libksg_oracle", text, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_c_abi_example_compiles_against_the_header_and_library(tmp_path):
    """examples/ksg_mi_cabi.cu uses nothing but include/ksg_b200.h, the CUDA runtime and
    libksg_b200.so (no torch, no Python): it must compile and link for sm_100a."""
    import shutil
    import subprocess

    from syntropy_b200 import _build

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    lib = _build.build()
    out = tmp_path / "ksg_mi_cabi"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"),
           "-o", str(out), os.path.join(ROOT, "examples", "ksg_mi_cabi.cu"), "-L", os.path.dirname(lib), "-lksg_b200"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    assert out.exists()
