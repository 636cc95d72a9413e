#!/bin/bash
# round 2: the ncu evidence committed under profiles/ (each capture after a plain run of the same command exited 0)
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
# 1. headline bench: launch list of the same command line
B="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-secondary --no-verify"
timeout 600 $B > gpurun_out/plain_bench.log 2>&1 && timeout 900 $L -c 400 --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "bench launch list rc $?"
# 2. the dominant kernel of the headline: the tcgen05 GEMM main pass at configs[3] (10M x 768, batch 256)
C="python tools/gemm_probe.py --once"
timeout 300 $C > gpurun_out/plain_g.log 2>&1 && timeout 900 $N -k regex:cosine_gemm_kernel -s 3 -c 1 -o gpurun_out/r02_prof_gemm $C > gpurun_out/ncu_g.log 2>&1
echo "gemm rc $?"
timeout 300 $C > /dev/null 2>&1 && timeout 600 $L -c 40 --log-file gpurun_out/r02_launches_gemm.csv $C > /dev/null 2>&1
# 3. BM25 (configs[2]: 10M docs, batch 1024), the shipped <640,2048> variant
C="python tools/bm25_probe.py --once --batch 1024"
timeout 300 $C > gpurun_out/plain_b.log 2>&1 && timeout 900 $N -k regex:bm25_blocked -s 2 -c 1 -o gpurun_out/r02_prof_bm25 $C > gpurun_out/ncu_b.log 2>&1
echo "bm25 rc $?"
# 4. the hybrid step at the 8-GPU shard size: per-kernel launch list
C="python tools/hybrid_probe.py --once"
timeout 300 $C > /dev/null 2>&1 && timeout 600 $L -c 60 --log-file gpurun_out/r02_launches_hybrid_6m.csv $C > /dev/null 2>&1
echo "hybrid launch list rc $?"
ls -la gpurun_out | grep r02_
