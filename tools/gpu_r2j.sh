#!/bin/bash
# Round-2 GPU session J: lazy axes (bucketed 1-D counts instead of the coordinate sorts), roomy lists A/B.
set -u
O=gpurun_out/r2j
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py -q -x ) > $O/pytest_fused.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fused.log
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
KSG_B200_WU_ROOMY=1 $B > $O/bench_roomy.json 2> $O/bench_roomy.err
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fused.py ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 22 ) > $O/fuzz.log 2>&1
( time timeout 300 python tools/quantised.py ) > $O/quantised.txt 2>&1
timeout 300 python tools/timeline.py > $O/timeline.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'axis_count_bucket|bucket_scatter|curve_key|knn_plan' -c 4 \
    -o $O/c2_buckets python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
echo done > $O/done.txt
