#!/bin/bash
# Round-2 GPU session Z (1 GPU): finer units for heavy lists (on top of the per-quarter merge): slices of an 8-rank
# step and the full step for fine_div = 1 (off), 2, 4.
set -u
O=gpurun_out/r2z
mkdir -p $O
rm -f $O/slices.txt
for f in 1 2 4; do
  echo "== KSG_B200_FINE_UNITS=$f" >> $O/slices.txt
  for r in 0 3 4 7; do
    KSG_B200_FINE_UNITS=$f python tools/slice_step.py --world 8 --rank $r >> $O/slices.txt 2>&1
  done
  KSG_B200_FINE_UNITS=$f python tools/slice_step.py --world 1 >> $O/slices.txt 2>&1
  KSG_B200_FINE_UNITS=$f python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline > $O/bench_fine$f.json 2> $O/bench_fine$f.err
done
echo done > $O/done.txt
