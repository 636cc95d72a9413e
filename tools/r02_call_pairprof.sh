#!/bin/bash
# CTA-pair GEMM kernel (opt-in): correctness, A/B against the default kernel on the same box, ncu capture
set -u
mkdir -p gpurun_out
timeout 600 python tools/pair_check.py > gpurun_out/pair_check.log 2>&1; echo "pair_check rc $?"; tail -3 gpurun_out/pair_check.log
for rep in 1 2; do
  for pr in 0 1; do
    OI_PAIR=$pr timeout 300 python tools/gemm_probe.py >> gpurun_out/pair_ab.log 2>&1
  done
done
grep -o '"default": \[[^]]*\]' gpurun_out/pair_ab.log
export OI_PAIR=1
C="python tools/gemm_probe.py --once"
timeout 300 $C > gpurun_out/plain_pair.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:cosine_gemm_pair -s 3 -c 1 -o gpurun_out/r02_prof_gemm_pair $C > gpurun_out/ncu_pair.log 2>&1
echo "pair ncu rc $?"
