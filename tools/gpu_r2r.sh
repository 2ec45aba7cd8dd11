#!/bin/bash
# Round-2 GPU session R (1 GPU): leave-one-out count shared between the ALU and FMA pipes (pair maxima +
# per-coordinate indicators): parity, C4 timing, ncu of the kernel.
set -u
O=gpurun_out/r2r
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q -x -k "loo or c4 or fused or multivariate or information or correlation" ) > $O/pytest_loo.log 2>&1
echo "pytest rc=$?" >> $O/pytest_loo.log
( time python bench.py --steps 10 --warmup 5 --configs c4 --no-dense-leg --no-cpu-baseline ) > $O/bench_c4.json 2> $O/bench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'count_f32_kernel' -c 1 \
    -o $O/c4_loo python tools/run_configs.py --configs c4 --check 64 --reps 1 > $O/ncu_loo.log 2>&1
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 31 ) > $O/fuzz31.log 2>&1
echo done > $O/done.txt
