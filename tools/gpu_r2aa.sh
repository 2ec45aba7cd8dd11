#!/bin/bash
# Round-2 GPU session AA (1 GPU): final code -- full GPU suite, fuzz, the driver's default bench line (timed), the
# reference arm, launch list and ncu of the final C2 kernels.
set -u
O=gpurun_out/r2aa
mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q -x ) > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
for s in 51 52; do
  ( time timeout 900 python tools/fuzz_engines.py --cases 200 --seed $s ) > $O/fuzz$s.log 2>&1
done
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
( time python bench.py --impl reference --steps 3 --warmup 3 ) > $O/bench_ref.json 2> $O/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_wunits_kernel|knn_plan_kernel|axis_count|finish_kernel' -c 4 \
    -o $O/c2_final python bench.py --steps 1 --warmup 3 --configs none --no-dense-leg --no-cpu-baseline > $O/ncu_full.log 2>&1
echo done > $O/done.txt
