#!/bin/bash
# A/B: the round-1 build against the current one on the same box (the tcgen05 cosine call alone)
set -u
mkdir -p gpurun_out
for i in 1 2; do
echo "== r1 build"; OI_GPU_LIB=$PWD/tools/probes/r1/libopenintel_gpu.so timeout 300 python tools/gemm_probe.py 2>gpurun_out/ab.err | tee gpurun_out/ab_r1_$i.json
echo "== current build"; timeout 300 python tools/gemm_probe.py 2>>gpurun_out/ab.err | tee gpurun_out/ab_cur_$i.json
done
tail -3 gpurun_out/ab.err
