#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gemm"; timeout 600 python -m pytest tests/test_gpu_cosine_gemm.py -m gpu -x -q > gpurun_out/pytest6.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest6.log
for i in 1 2; do
echo "== current build"; timeout 300 python tools/gemm_probe.py 2>>gpurun_out/ab.err | tee gpurun_out/ab_cur_$i.json
echo "== r1 build"; OI_GPU_LIB=$PWD/tools/probes/r1/libopenintel_gpu.so timeout 300 python tools/gemm_probe.py 2>gpurun_out/ab.err | tee gpurun_out/ab_r1_$i.json
done
echo "== launch list"
timeout 300 python tools/gemm_probe.py --once > gpurun_out/plain_gemm.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_gemm.csv python tools/gemm_probe.py --once > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; grep -v "^==" gpurun_out/launches_gemm.csv | awk -F'","' '{print substr($5,1,60), $(NF)}' | tail -14
