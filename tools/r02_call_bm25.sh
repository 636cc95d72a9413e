#!/bin/bash
# BM25 kernel iteration: parity + A/B against the committed build on the same box
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_full_size_oracle.py tests/test_gpu_store.py -x -q -m gpu > gpurun_out/bm25_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/bm25_tests.log
tail -3 gpurun_out/bm25_tests.log
rm -f gpurun_out/bm25_sweep.log
for rep in 1; do
for v in base intree; do
  if [ $v = intree ]; then unset OI_GPU_LIB; else export OI_GPU_LIB=$GRAFT_REPO_ROOT/tools/probes/$v/libopenintel_gpu.so; fi
  echo "== $v" >> gpurun_out/bm25_sweep.log
  timeout 600 python tools/bm25_sweep.py --docs 6250000 >> gpurun_out/bm25_sweep.log 2>&1
  timeout 600 python tools/bm25_sweep.py --docs 50000000 >> gpurun_out/bm25_sweep.log 2>&1
  timeout 600 python tools/bm25_sweep.py --docs 10000000 --batch 1024 >> gpurun_out/bm25_sweep.log 2>&1
  timeout 600 python tools/bm25_sweep.py --docs 1000000 --batch 1 >> gpurun_out/bm25_sweep.log 2>&1; timeout 600 python tools/bm25_sweep.py --docs 12500000 >> gpurun_out/bm25_sweep.log 2>&1
done
done
unset OI_GPU_LIB
cat gpurun_out/bm25_sweep.log
