#!/bin/bash
# CTA-pair GEMM: ring depth A/B (48 slabs = 192 KB per CTA vs 24 = 96 KB), alternating processes on one box
set -u
mkdir -p gpurun_out
rm -f gpurun_out/ring_ab.log
for rep in 1 2; do
  for ring in 48 24; do
    OI_PAIR=1 OI_PAIR_RING=$ring timeout 300 python tools/gemm_probe.py 2>&1 | grep -o '"default": \[[^]]*\]' | sed "s/^/ring $ring /" | tee -a gpurun_out/ring_ab.log
  done
done
