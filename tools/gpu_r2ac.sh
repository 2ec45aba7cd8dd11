#!/bin/bash
# Round-2 GPU session AC (1 GPU): grouped count kernel (8-query groups in the own-order marginal passes):
# parity, fuzz, C3 / C5 with and without it.
set -u
O=gpurun_out/r2ac
mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 61 ) > $O/fuzz61.log 2>&1
python bench.py --steps 3 --warmup 3 --configs c3,c5 --no-dense-leg --no-cpu-baseline > $O/bench_grouped.json 2> $O/bench_grouped.err
KSG_B200_COUNT_GROUPS=0 python bench.py --steps 3 --warmup 3 --configs c3,c5 --no-dense-leg --no-cpu-baseline > $O/bench_whole_warp.json 2> $O/bench_whole_warp.err
( time timeout 300 python tools/run_temporal.py ) > $O/temporal.txt 2>&1
echo done > $O/done.txt
