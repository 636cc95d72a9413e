#!/bin/bash
# how much of the BM25 launch on a small shard is tail (SMs idle at the end)?  active vs elapsed cycles per SM
set -u
mkdir -p gpurun_out
C="python tools/bm25_probe.py --once --batch 256 --docs 6250000"
timeout 300 $C > /dev/null 2>&1 && timeout 600 ncu --metrics sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,sm__inst_executed.avg.per_cycle_elapsed,sm__inst_executed.avg.per_cycle_active,gpu__time_duration.sum,smsp__warps_active.avg.per_cycle_active --clock-control none -k regex:bm25_blocked -s 2 -c 1 --csv $C 2>/dev/null | grep -v "^==" | cut -d, -f5,13- | tail -12
