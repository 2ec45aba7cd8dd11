#!/bin/bash
# Round-2 GPU session T (1 GPU): width of the curve key (radix passes of the key sort against box tightness).
set -u
O=gpurun_out/r2t
mkdir -p $O
B="python bench.py --steps 20 --warmup 5 --configs none --no-dense-leg --no-cpu-baseline"
$B > $O/bench_key32.json 2> $O/bench_key32.err
for kb in 28 24 20 16; do
  KSG_B200_KEY_BITS=$kb $B > $O/bench_key$kb.json 2> $O/bench_key$kb.err
done
$B > $O/bench_key32_b.json 2> $O/bench_key32_b.err
echo done > $O/done.txt
