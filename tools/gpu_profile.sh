#!/bin/bash
# tools/gpu_profile.sh — one gpurun call: ncu --set full captures of the dominant kernels (scan, tcgen05 GEMM, BM25, multi-query scan) (each after a
# plain run of the same command exited 0), plus launch lists.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
C1="python bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 3"
timeout 300 $C1 > gpurun_out/p_scan_plain.log 2>&1 && timeout 600 $N -k regex:cosine_scan -s 3 -c 1 -o gpurun_out/prof_scan $C1 > gpurun_out/p_scan_ncu.log 2>&1
echo "scan rc $?"
timeout 300 $C1 > /dev/null 2>&1 && timeout 600 $L -c 60 --log-file gpurun_out/launches_scan.csv $C1 > /dev/null 2>&1
C2="python tools/bench_workloads.py gemm --batch 256 --steps 2 --warmup 2"
timeout 300 $C2 > gpurun_out/p_gemm_plain.log 2>&1 && timeout 600 $N -k regex:cosine_gemm_kernel -s 14 -c 1 -o gpurun_out/prof_gemm $C2 > gpurun_out/p_gemm_ncu.log 2>&1
echo "gemm rc $?"
timeout 300 $C2 > /dev/null 2>&1 && timeout 600 $L -c 80 --log-file gpurun_out/launches_gemm.csv $C2 > /dev/null 2>&1
C3="python tools/bench_workloads.py bm25 --batch 1024 --steps 2 --warmup 2"
timeout 300 $C3 > gpurun_out/p_bm25_plain.log 2>&1 && timeout 900 $N -k regex:bm25_blocked -s 2 -c 1 -o gpurun_out/prof_bm25 $C3 > gpurun_out/p_bm25_ncu.log 2>&1
echo "bm25 rc $?"
timeout 300 $C3 > /dev/null 2>&1 && timeout 600 $L -c 60 --log-file gpurun_out/launches_bm25.csv $C3 > /dev/null 2>&1
C4="python tools/multi_scan_probe.py"
timeout 300 $C4 > gpurun_out/p_multi_plain.log 2>&1 && timeout 600 $N -k regex:cosine_scan_bulk_multi -s 4 -c 1 -o gpurun_out/prof_bmulti $C4 > gpurun_out/p_multi_ncu.log 2>&1
echo "multi-query scan rc $?"
ls -la gpurun_out | grep -E "prof_|launches_"
