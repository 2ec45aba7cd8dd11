"""CPU check of axis_count_bucket_kernel (csrc/axis_buckets.cuh), the exact 1-D counts of prepared sets
with lazy axes (no coordinate sort: value buckets + prefix sums + the exact predicate at the two ends):
the kernel and the curve-position functions are cut out of the CUDA sources, compiled for the host
(tests/host_shims/bucket_count_harness.cpp, directed rounding through <cfenv>) and compared with brute
force on normal, quantised, heavy-tailed, constant, offset and few-valued data with radii of every kind."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut_sources():
    prep = open(os.path.join(ROOT, "syntropy_b200", "csrc", "prepare.cuh")).read()
    a = prep.index("// position of the centred value x among the CURVE_TABLE ascending table values")
    b = prep.index("// keys[i] = Hilbert index of point i's curve coordinates.")
    curve = prep[a:b]
    src = open(os.path.join(ROOT, "syntropy_b200", "csrc", "axis_buckets.cuh")).read()
    k0 = src.index("constexpr int BUCKET_SCAN_CAP")
    kernel = src[k0:src.index("}  // namespace ksg", k0)]
    assert "axis_count_bucket_kernel" in kernel and "curve_pos" in curve and "bucket_bits" in curve
    return curve + "\n" + kernel


def test_bucket_count_kernel_against_brute_force_on_the_host(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    (tmp_path / "bucket_kernel.inc").write_text(cut_sources())
    shutil.copy(os.path.join(ROOT, "tests", "host_shims", "bucket_count_harness.cpp"), tmp_path / "bucket_count_harness.cpp")
    exe = tmp_path / "bucket_count_harness"
    proc = subprocess.run([gxx, "-O1", "-std=c++17", "-frounding-math", "-o", str(exe), "bucket_count_harness.cpp"],
                          cwd=tmp_path, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    run = subprocess.run([str(exe), "120"], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-3000:]
    words = run.stdout.split()
    stats = dict(zip(words[0::2], map(int, words[1::2])))
    assert stats["wrong"] == 0 and stats["checked"] > 200_000
    assert stats["flagged_runs"] > 0 and stats["heavy_runs"] > 0   # the give-up paths were exercised too
