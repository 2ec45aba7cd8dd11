#!/bin/bash
# Round-2 GPU session AH (1 GPU): final code -- full GPU suite, fuzz (one seed with the grouped LOO kernel forced on),
# the driver's default bench line, smoke, launch list of the C2 step, ncu of the grouped count kernels.
set -u
O=gpurun_out/r2ah
mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
( time timeout 900 python tools/fuzz_engines.py --cases 200 --seed 81 ) > $O/fuzz81.log 2>&1
( time KSG_B200_LOO_GROUPS=force timeout 900 python tools/fuzz_engines.py --cases 200 --seed 82 ) > $O/fuzz82_forced.log 2>&1
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'count_loo_group_kernel' -c 1 \
    -o $O/c4_loo_group python tools/run_configs.py --configs c4 --check 64 --reps 1 > $O/ncu_loo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'count_gunits_kernel' -c 3 \
    -o $O/c3_count_group python tools/run_configs.py --configs c3 --check 64 --reps 1 > $O/ncu_c3.log 2>&1
echo done > $O/done.txt
