"""CPU checks of two device functions that are plain C++ (compiled for the host with g++,
tests/host_shims/curve_psi_harness.cpp): the Hilbert index behind the prepared order is a bijection
whose consecutive keys are neighbouring cells (2 to 8 dimensions), and the device digamma equals
scipy's ``digamma`` (the reference's psi, tests/golden/digamma.npz) on every integer argument."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut_device_functions():
    prep = open(os.path.join(ROOT, "syntropy_b200", "csrc", "prepare.cuh")).read()
    h0 = prep.index("template <int CAP>\n__device__ __forceinline__ unsigned long long hilbert_index")
    hilbert = prep[h0:prep.index("\n}\n", h0) + 3]  # the function body only (kernels follow it)
    epi = open(os.path.join(ROOT, "syntropy_b200", "csrc", "epilogue.cuh")).read()
    p0 = epi.index("__device__ inline double psi_of_int")
    psi = epi[p0:epi.index("static __global__ void psi_table_kernel")]
    return hilbert + "\n" + psi


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    tmp = tmp_path_factory.mktemp("curve_psi")
    (tmp / "device_functions.inc").write_text(cut_device_functions())
    shutil.copy(os.path.join(ROOT, "tests", "host_shims", "curve_psi_harness.cpp"), tmp / "curve_psi_harness.cpp")
    exe = tmp / "curve_psi_harness"
    proc = subprocess.run([gxx, "-O1", "-std=c++17", "-o", str(exe), "curve_psi_harness.cpp"], cwd=tmp,
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    return str(exe)


def test_hilbert_index_is_a_bijection_with_unit_steps(harness):
    run = subprocess.run([harness, "curve"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and run.stdout.split() == ["curve_violations", "0"], run.stdout


def test_device_digamma_equals_scipy(harness, golden):
    g = golden("digamma")
    args = g["args"].astype(np.int64)
    run = subprocess.run([harness], input="\n".join(str(int(a)) for a in args), capture_output=True, text=True,
                         timeout=600)
    assert run.returncode == 0
    mine = np.array([float(t) for t in run.stdout.split()])
    ref = g["psi"]
    assert mine.shape == ref.shape
    assert np.isneginf(mine[0]) and np.isneginf(ref[0])
    # harmonic range: bit-exact; asymptotic range: the series is cephes', only log() may differ by an ulp
    small = (args >= 1) & (args <= 10)
    assert np.array_equal(mine[small], ref[small])
    big = args > 10
    assert np.all(np.abs(mine[big] - ref[big]) <= 4e-16 * np.abs(ref[big]))
