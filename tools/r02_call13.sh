#!/bin/bash
set -u
mkdir -p gpurun_out
for docs in 6250000 25000000; do
for ipw in 2 4 8 16 24; do
  echo "== docs $docs ipw $ipw"; OI_IPW=$ipw timeout 200 python tools/bm25_probe.py --docs $docs --batch 256 2>>gpurun_out/ipw.err | tee -a gpurun_out/ipw.jsonl
done; done
tail -2 gpurun_out/ipw.err
