#!/bin/bash
set -u
mkdir -p gpurun_out
C="python tools/bm25_probe.py --once --no-defer 1"
timeout 300 $C > gpurun_out/plain_b.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_blocked -s 2 -c 1 -f -o gpurun_out/prof_bm25_sweep $C > gpurun_out/ncu_b.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_b.log
C="python tools/bm25_probe.py --once --no-defer 0"
timeout 300 $C > gpurun_out/plain_b2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_blocked -s 2 -c 1 -f -o gpurun_out/prof_bm25_defer $C > gpurun_out/ncu_b2.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_b2.log
