#!/bin/bash
# Round-2 GPU session I: launch fusions (gather tile kernel, multi-axis counts, fused tail), warp-level units
# kernel with compile-time list capacity and the unsorted refine: parity, A/B, fuzz, ncu.
set -u
O=gpurun_out/r2i
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py -q -x ) > $O/pytest_fused.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fused.log
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
KSG_B200_UNITS=cta $B > $O/bench_cta.json 2> $O/bench_cta.err
KSG_B200_FUSED_GATHER=0 $B > $O/bench_unfused_gather.json 2> $O/bench_unfused_gather.err
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fused.py ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 21 ) > $O/fuzz.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_wunits_kernel|gather_tile|finish_kernel|axis_count' -c 4 \
    -o $O/c2_new python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
echo done > $O/done.txt
