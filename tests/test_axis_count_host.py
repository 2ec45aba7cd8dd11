"""CPU check of axis_count_kernel (csrc/fast_count.cuh), the 1-D marginal counts of the path: cut out
of the header, compiled for the host (tests/host_shims/axis_count_harness.cpp) and run block by
block against brute force -- array lengths from 1 to 6000 (shorter and longer than the 1024-entry
sample table), ties, constant arrays, heavy tails, radii 0 / tiny / inf, strict and closed balls,
query slices that do not start at 0."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut_axis_count_kernel():
    src = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_count.cuh")).read()
    a = src.index("constexpr int AXIS_THREADS")
    return src[a:src.index("static __global__ void scatter_f64_kernel", a)]


def test_axis_count_kernel_against_brute_force_on_the_host(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    (tmp_path / "axis_count_kernel.inc").write_text(cut_axis_count_kernel())
    shutil.copy(os.path.join(ROOT, "tests", "host_shims", "axis_count_harness.cpp"), tmp_path / "axis_count_harness.cpp")
    exe = tmp_path / "axis_count_harness"
    flags = ["-DHAVE_SORTED_KERNEL"] if "axis_count_sorted_kernel" in cut_axis_count_kernel() else []
    proc = subprocess.run([gxx, "-O1", "-std=c++17", *flags, "-o", str(exe), "axis_count_harness.cpp"], cwd=tmp_path,
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    run = subprocess.run([str(exe), "300"], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-3000:]
    words = run.stdout.split()
    stats = dict(zip(words[0::2], map(int, words[1::2])))
    assert stats["wrong"] == 0 and stats["checked"] > 200_000
