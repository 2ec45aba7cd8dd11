#!/bin/bash
# Round-2 GPU session L: grouped k-NN kernel v2 (owner lanes + hit bits), direct launches vs graphs, fuzz details.
set -u
O=gpurun_out/r2l
mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_fused.py -q -x ) > $O/pytest_fused.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fused.log
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
KSG_B200_GRAPHS=0 $B > $O/bench_nographs.json 2> $O/bench_nographs.err
( time python bench.py --steps 5 --warmup 3 --configs c3,c4,c5 --no-dense-leg --no-cpu-baseline ) > $O/bench_cfg_grouped.json 2> $O/bench_cfg_grouped.err
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 23 ) > $O/fuzz23.log 2>&1
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 21 ) > $O/fuzz21.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fused.py ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_gunits' -c 1 \
    -o $O/c3_gunits python bench.py --steps 1 --warmup 1 --configs c3 --no-dense-leg --no-cpu-baseline > $O/ncu_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
echo done > $O/done.txt
