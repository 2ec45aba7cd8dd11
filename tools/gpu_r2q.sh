#!/bin/bash
# Round-2 GPU session Q (1 GPU): leave-one-out count by indicator sums (no maxima): parity, C4 timing,
# launch list + ncu of the round's final C2 kernels and of the LOO count kernel.
set -u
O=gpurun_out/r2q
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q -x -k "loo or c4 or fused or multivariate or information or correlation" ) > $O/pytest_loo.log 2>&1
echo "pytest rc=$?" >> $O/pytest_loo.log
( time python bench.py --steps 10 --warmup 5 --configs c4 --no-dense-leg --no-cpu-baseline ) > $O/bench_c4.json 2> $O/bench_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_wunits_kernel|knn_plan_kernel|axis_count|finish_kernel|curve_key|gather_tile' -c 8 \
    -o $O/c2_final python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'count_f32_kernel' -c 1 \
    -o $O/c4_loo python tools/run_configs.py --configs c4 --check 64 --reps 1 > $O/ncu_loo.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
echo done > $O/done.txt
