#!/bin/bash
# Round-2 GPU session X (1 GPU): one rank's share of the sharded C2 step (1/8, 1/2 of the queries), spans + launch
# list; finer units for heavy lists: parity + fuzz.
set -u
O=gpurun_out/r2x
mkdir -p $O
python tools/slice_step.py --world 8 > $O/slice8_fine.txt 2>&1
python tools/slice_step.py --world 2 > $O/slice2_fine.txt 2>&1
python tools/slice_step.py --world 1 > $O/slice1_fine.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_slice8_fine.csv \
    python tools/slice_step.py --world 8 --steps 2 > $O/ncu_slice8.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 41 ) > $O/fuzz41.log 2>&1
python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
echo done > $O/done.txt
