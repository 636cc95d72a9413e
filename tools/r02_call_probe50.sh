#!/bin/bash
# probe-pass size at the headline shard (50M rows): default (1/256 of the tiles) vs 4 and 32 tiles per CTA
set -u
mkdir -p gpurun_out
timeout 600 python tools/gemm_probe.py --docs 50000000 2>&1 | tail -1 | tee gpurun_out/gemm_probe_50m.log
