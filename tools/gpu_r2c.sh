#!/bin/bash
# Round-2 GPU session C: Minkowski / drop-in / sort tests, band-only refine + heavy-first units A/B.
set -u
O=gpurun_out/r2c
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
KSG_B200_FUSED_REFINE=0 $B > $O/bench_unfused.json 2> $O/bench_unfused.err
( time timeout 600 python tools/fuzz_engines.py --cases 150 --seed 12 ) > $O/fuzz.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_units_kernel|knn_plan_kernel' -c 2 \
    -o $O/c2_knn python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
echo done > $O/done.txt
