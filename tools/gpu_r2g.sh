#!/bin/bash
# Round-2 GPU session G: default bench line twice (e2e per call), quantised data, final ncu captures.
set -u
O=gpurun_out/r2g
mkdir -p $O
( time python bench.py ) > $O/bench_default_1.json 2> $O/bench_default_1.err
( time python bench.py ) > $O/bench_default_2.json 2> $O/bench_default_2.err
python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline > $O/bench_quick.json 2> $O/bench_quick.err
( time python bench.py --impl reference --steps 3 --warmup 3 ) > $O/bench_ref.json 2> $O/bench_ref.err
( time timeout 300 python tools/quantised.py ) > $O/quantised.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_units_kernel|knn_plan_kernel|axis_count|psi_epilogue' -c 5 \
    -o $O/c2_final python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
echo done > $O/done.txt
