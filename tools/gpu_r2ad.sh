#!/bin/bash
# Round-2 GPU session AD (1 GPU): count kernel with eight groups of 4 queries against four groups of 8.
set -u
O=gpurun_out/r2ad
mkdir -p $O
for g in 8 4; do
  KSG_B200_COUNT_GROUPS=$g python bench.py --steps 3 --warmup 3 --configs c3,c5 --no-dense-leg --no-cpu-baseline > $O/bench_groups$g.json 2> $O/bench_groups$g.err
done
( time KSG_B200_COUNT_GROUPS=8 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "windowed or full_size or c3 or c5" ) > $O/pytest_groups8.log 2>&1
echo "pytest rc=$?" >> $O/pytest_groups8.log
echo done > $O/done.txt
