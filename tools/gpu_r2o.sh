#!/bin/bash
# Round-2 GPU session O: side sorts on the prepared copy (lifetime race of the caller's array fixed): full suite, fuzz x4.
set -u
O=gpurun_out/r2o
mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
for s in 21 23 24 25; do
  ( time timeout 900 python tools/fuzz_engines.py --cases 200 --seed $s ) > $O/fuzz$s.log 2>&1
done
( time timeout 300 python tools/fuzz_engines.py --cases 200 --seed 24 --only 195 --repeat 12 ) > $O/fuzz24_case195_x12.log 2>&1
( time python bench.py --steps 20 --warmup 5 --no-dense-leg --no-cpu-baseline ) > $O/bench_full.json 2> $O/bench_full.err
echo done > $O/done.txt
