#!/bin/bash
# Round-2 GPU session E: early order (delayed side sorts), exact-fp32 mode: parity, A/B, timeline, sanitizer.
set -u
O=gpurun_out/r2e
mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
KSG_B200_EARLY_ORDER=0 $B > $O/bench_exact_order.json 2> $O/bench_exact_order.err
KSG_B200_BINDING=ctypes $B > $O/bench_ctypes.json 2> $O/bench_ctypes.err
( time timeout 600 python tools/fuzz_engines.py --cases 200 --seed 14 ) > $O/fuzz.log 2>&1
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_full.json 2> $O/bench_full.err
timeout 300 python tools/timeline.py > $O/timeline.txt 2>&1
timeout 300 python tools/timeline.py --e2e > $O/timeline_e2e.txt 2>&1
( time timeout 300 python tools/quantised.py ) > $O/quantised.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 1500 bash tools/gpu_sanitize.sh > $O/sanitize_driver.log 2>&1
echo done > $O/done.txt
