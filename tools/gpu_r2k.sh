#!/bin/bash
# Round-2 GPU session K: finer value buckets + two-launch scan; grouped (8-query) k-NN kernel for 5..8-D: parity, A/B.
set -u
O=gpurun_out/r2k
mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_fused.py -q -x ) > $O/pytest_fused.log 2>&1
echo "pytest rc=$?" >> $O/pytest_fused.log
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
python -c "
import os
os.environ['X']='1'
from syntropy_b200 import _engine as E
" 
( time python bench.py --steps 5 --warmup 3 --configs c3,c4,c5 --no-dense-leg --no-cpu-baseline ) > $O/bench_cfg_grouped.json 2> $O/bench_cfg_grouped.err
( time KSG_B200_GROUPS=0 python bench.py --steps 5 --warmup 3 --configs c3,c4,c5 --no-dense-leg --no-cpu-baseline ) > $O/bench_cfg_warp.json 2> $O/bench_cfg_warp.err
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fused.py ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 23 ) > $O/fuzz.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'axis_count_bucket|bucket_scatter|knn_plan' -c 3 \
    -o $O/c2_buckets python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
echo done > $O/done.txt
