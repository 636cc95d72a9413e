#!/bin/bash
# launch lists of the final build (ncu --metrics gpu__time_duration.sum), each after a plain run of the same command
set -u
mkdir -p gpurun_out
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
B="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-secondary --no-verify"
timeout 600 $B > gpurun_out/plain_bench.log 2>&1 && timeout 900 $L -c 400 --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "bench launch list rc $?"
C="python tools/gemm_probe.py --once"
timeout 300 $C > /dev/null 2>&1 && timeout 600 $L -c 40 --log-file gpurun_out/r02_launches_gemm.csv $C > /dev/null 2>&1
echo "gemm launch list rc $?"
C="python tools/hybrid_probe.py --once"
timeout 300 $C > /dev/null 2>&1 && timeout 600 $L -c 60 --log-file gpurun_out/r02_launches_hybrid_6m.csv $C > /dev/null 2>&1
echo "hybrid launch list rc $?"
