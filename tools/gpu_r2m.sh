#!/bin/bash
# Round-2 GPU session M: grouped kernel v3 (the copy warp skips tiles no group reaches), fuzz repro with details.
set -u
O=gpurun_out/r2m
mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_fused.py -q -x ) > $O/pytest_fused.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fused.log
( time python bench.py --steps 5 --warmup 3 --configs c3,c4,c5 --no-dense-leg --no-cpu-baseline ) > $O/bench_cfg.json 2> $O/bench_cfg.err
( time timeout 300 python tools/fuzz_engines.py --cases 200 --seed 23 --only 73 --repeat 4 ) > $O/fuzz23_case73.log 2>&1
( time timeout 300 python tools/fuzz_engines.py --cases 200 --seed 21 --only 170 --repeat 4 ) > $O/fuzz21_case170.log 2>&1
( time KSG_B200_GROUPS=0 timeout 300 python tools/fuzz_engines.py --cases 200 --seed 21 --only 170 --repeat 2 ) > $O/fuzz21_case170_nogroups.log 2>&1
( time timeout 900 python tools/fuzz_engines.py --cases 200 --seed 24 ) > $O/fuzz24.log 2>&1
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_gunits' -c 1 \
    -o $O/c3_gunits python bench.py --steps 1 --warmup 1 --configs c3 --no-dense-leg --no-cpu-baseline > $O/ncu_c3.log 2>&1
echo done > $O/done.txt
