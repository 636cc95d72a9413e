#!/bin/bash
set -u
mkdir -p gpurun_out
C="python tools/gemm_probe.py --once"
timeout 300 $C > gpurun_out/plain_g.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cosine_gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_main $C > gpurun_out/ncu_g.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_g.log
C="python tools/gemm_probe.py --once --debug 1"
timeout 300 $C > gpurun_out/plain_g1.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:cosine_gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_tmaonly $C > gpurun_out/ncu_g1.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_g1.log
C="python tools/gemm_probe.py --once --batch 128"
timeout 300 $C > gpurun_out/plain_g128.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:cosine_gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_b128 $C > gpurun_out/ncu_g128.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/ncu_g128.log
