#!/bin/bash
set -u
mkdir -p gpurun_out
C="python tools/gemm_probe.py --once --debug 8"
timeout 300 $C > gpurun_out/plain_g8.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cosine_gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_dbg8 $C > gpurun_out/ncu_g8.log 2>&1
echo "ncu rc $?"; tail -3 gpurun_out/ncu_g8.log
echo "== bm25 tests"; timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_full_size_oracle.py -m gpu -x -q > gpurun_out/pytest5.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest5.log
echo "== bm25 probe"; timeout 600 python tools/bm25_probe.py 2>gpurun_out/bm25_probe.err | tee gpurun_out/bm25_probe.json; tail -3 gpurun_out/bm25_probe.err
echo "== bm25 probe r1"; OI_GPU_LIB=$PWD/tools/probes/r1/libopenintel_gpu.so timeout 600 python tools/bm25_probe.py 2>>gpurun_out/bm25_probe.err | tee gpurun_out/bm25_probe_r1.json
