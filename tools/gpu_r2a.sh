#!/bin/bash
# Round-2 GPU session A: parity suite, A/B of the windowed k-NN switches, full bench line, launch list.
set -u
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_new.json 2> $O/bench_new.err
KSG_B200_TIGHT_SEED=0 $B > $O/bench_legacyseed.json 2> $O/bench_legacyseed.err
KSG_B200_FUSED_REFINE=0 $B > $O/bench_unfused.json 2> $O/bench_unfused.err
KSG_B200_TIGHT_SEED=0 KSG_B200_FUSED_REFINE=0 $B > $O/bench_legacy.json 2> $O/bench_legacy.err
( time python bench.py --steps 20 --warmup 5 ) > $O/bench_full.json 2> $O/bench_full.err
( time python bench.py --impl reference --steps 3 --warmup 3 ) > $O/bench_ref.json 2> $O/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
echo done > $O/done.txt
