#!/bin/bash
# round 2, GPU call 2: the GEMM with the in-pass threshold ladder: parity, timing per probe size, per-kernel breakdown.
set -u
mkdir -p gpurun_out
echo "== pytest gemm + bm25 + full-size"; timeout 900 python -m pytest tests/test_gpu_cosine_gemm.py tests/test_gpu_bm25.py tests/test_gpu_full_size_oracle.py -m gpu -x -q > gpurun_out/pytest2.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest2.log
echo "== gemm probe"; timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe.json 2> gpurun_out/gemm_probe.err; echo "exit $?"; cat gpurun_out/gemm_probe.json; tail -3 gpurun_out/gemm_probe.err
echo "== gemm probe lite"; timeout 300 python tools/gemm_probe.py --lite 1 > gpurun_out/gemm_probe_lite.json 2>> gpurun_out/gemm_probe.err; echo "exit $?"; cat gpurun_out/gemm_probe_lite.json
echo "== launch list"
timeout 300 python tools/gemm_probe.py --once > gpurun_out/plain_gemm.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_gemm.csv python tools/gemm_probe.py --once > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; grep -v "^==" gpurun_out/launches_gemm.csv | awk -F'","' '{print $5, $(NF)}' | tail -25
