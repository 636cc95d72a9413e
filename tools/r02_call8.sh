#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pair check"; timeout 120 python tools/pair_check.py > gpurun_out/pair_check.log 2>&1; echo "exit $?"; tail -12 gpurun_out/pair_check.log
echo "== pytest gemm"; timeout 300 python -m pytest tests/test_gpu_cosine_gemm.py -m gpu -x -q > gpurun_out/pytest8.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest8.log
echo "== probe (pair ring 48)"; timeout 200 python tools/gemm_probe.py 2>gpurun_out/p8.err | tee gpurun_out/pair48.json
echo "== probe (pair ring 24)"; OI_PAIR_RING=24 timeout 200 python tools/gemm_probe.py 2>>gpurun_out/p8.err | tee gpurun_out/pair24.json
echo "== probe (no pair)"; OI_PAIR=0 timeout 200 python tools/gemm_probe.py 2>>gpurun_out/p8.err | tee gpurun_out/nopair.json
tail -3 gpurun_out/p8.err
