#!/bin/bash
# Round-2 GPU session H: warp-autonomous units kernel (DP <= 4): parity, fuzz, A/B against the CTA-level kernel, ncu.
set -u
O=gpurun_out/r2h
mkdir -p $O
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_warp.json 2> $O/bench_warp.err
KSG_B200_UNITS=cta $B > $O/bench_cta.json 2> $O/bench_cta.err
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 21 ) > $O/fuzz.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_wunits_kernel' -c 2 \
    -o $O/c2_wunits python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
( time python bench.py --steps 20 --warmup 5 --no-dense-leg --no-cpu-baseline ) > $O/bench_configs.json 2> $O/bench_configs.err
echo done > $O/done.txt
