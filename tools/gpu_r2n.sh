#!/bin/bash
set -u
O=gpurun_out/r2n
mkdir -p $O
timeout 300 python tools/repro_idx.py --seed 21 --case 170 --repeat 6 > $O/repro_grouped.log 2>&1
KSG_B200_GROUPS=0 timeout 300 python tools/repro_idx.py --seed 21 --case 170 --repeat 6 > $O/repro_nogroups.log 2>&1
timeout 300 python tools/repro_idx.py --seed 24 --case 196 --repeat 4 > $O/repro_24_196.log 2>&1
KSG_B200_FUSED_REFINE=0 timeout 300 python tools/repro_idx.py --seed 21 --case 170 --repeat 4 > $O/repro_unfused.log 2>&1
echo done > $O/done.txt
