#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gemm"; timeout 600 python -m pytest tests/test_gpu_cosine_gemm.py -m gpu -x -q > gpurun_out/pytest4.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest4.log
for i in 1 2; do
echo "== current build"; timeout 300 python tools/gemm_probe.py 2>>gpurun_out/ab.err | tee gpurun_out/ab_cur_$i.json
echo "== r1 build"; OI_GPU_LIB=$PWD/tools/probes/r1/libopenintel_gpu.so timeout 300 python tools/gemm_probe.py 2>gpurun_out/ab.err | tee gpurun_out/ab_r1_$i.json
done
tail -3 gpurun_out/ab.err
