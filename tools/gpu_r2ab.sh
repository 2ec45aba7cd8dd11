#!/bin/bash
# Round-2 GPU session AB (1 GPU): seed window of the planning pass in 6-D / 8-D (C3, C4, C5).
set -u
O=gpurun_out/r2ab
mkdir -p $O
for w in 768 1536; do
  KSG_B200_SEED_W=$w python bench.py --steps 3 --warmup 3 --configs c3,c4,c5 --no-dense-leg --no-cpu-baseline > $O/bench_seedw$w.json 2> $O/bench_seedw$w.err
done
echo done > $O/done.txt
