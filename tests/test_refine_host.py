"""CPU check of the LOGIC of knn_refine_kernel (csrc/fast_knn.cuh): its CUDA source is plain C++
apart from a few keywords, so it is cut out of the header, compiled for the host with g++
(tests/host_shims/refine_harness.cpp) and run one call per "thread" on emulated filter output --
disjoint candidate chunks, each keeping its kcap smallest fp32 distances.  Every query the kernel
decides must carry the exact float64 k'-th neighbour distance (and, when asked, the k' neighbours
ordered by (distance, original index)) of a brute-force search; continuous data must never be
flagged; tie-saturated data may be flagged but never answered wrongly."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut_refine_kernel():
    src = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_knn.cuh")).read()
    launch = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_launch.h")).read()
    args0 = launch.index("struct RefineArgs {")
    args = launch[args0:launch.index("};", args0) + 2]
    band0 = src.index("__host__ __device__ inline double band_query_f64")
    band = src[band0:src.index("}", band0) + 1]
    r0 = src.index("__host__ __device__ inline bool fp32_range_ok")
    band += "\n" + src[r0:src.index("\n}\n", r0) + 3]
    seed0 = src.index("__device__ __forceinline__ float tight_seed")
    seed = src[seed0:src.index("}", seed0) + 1]
    start = src.index("// Exact refine: see the file header.")
    kernel = src[start:src.index("}  // namespace ksg", start)]
    assert re.search(r"\bknn_refine_kernel\(", kernel) and re.search(r"\brefine_single_list\(", kernel)
    warp = open(os.path.join(ROOT, "syntropy_b200", "csrc", "fast_knn_warp.cuh")).read()
    c0 = warp.index("constexpr int WU_MAX_KC")
    consts = warp[c0:warp.index("\n", warp.index("constexpr int WU_BAND_REGS")) + 1]
    cap0 = warp.index("constexpr int wunits_list_capacity")
    consts += warp[cap0:warp.index("\n", cap0) + 1]
    u0 = warp.index("template <int DP, int KC, int KPR>")
    unsorted = warp[u0:warp.index("\n}\n", u0) + 3]
    assert "refine_unsorted_list" in unsorted
    return args + "\n\n" + band + "\n\n" + seed + "\n\n" + kernel + "\n\n" + consts + "\n" + unsorted


def test_refine_kernel_logic_on_the_host(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    (tmp_path / "refine_kernel.inc").write_text(cut_refine_kernel())
    shutil.copy(os.path.join(ROOT, "tests", "host_shims", "refine_harness.cpp"), tmp_path / "refine_harness.cpp")
    exe = tmp_path / "refine_harness"
    proc = subprocess.run([gxx, "-O1", "-std=c++17", "-DHAVE_UNSORTED", "-o", str(exe), "refine_harness.cpp"], cwd=tmp_path,
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    run = subprocess.run([str(exe), "250"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-3000:]
    stats = dict(zip(run.stdout.split()[0::2], map(int, run.stdout.split()[1::2])))
    assert stats["wrong"] == 0
    assert stats["checked"] > 30_000
    assert stats["flagged_continuous"] == 0       # the k' + 3 margin always suffices without ties
    assert stats["flagged_tied"] > 0 and stats["decided_tied"] > 0   # both outcomes were exercised
    # the fused refine of single-unit blocks (refine_single_list) under tight seeds
    assert stats["single_checked"] > 15_000
    assert stats["single_flagged_continuous"] == 0
    # the unsorted-list refine of the warp-level kernel (refine_unsorted_list): equal to the sorted path and to
    # brute force whenever it answers; continuous data practically never needs the general path
    assert stats["unsorted_checked"] > 15_000
    assert stats["unsorted_fallback_continuous"] <= stats["unsorted_checked"] // 50
    assert stats["unsorted_fallback_tied"] > 0
