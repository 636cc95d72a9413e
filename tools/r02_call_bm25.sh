#!/bin/bash
# BM25 after the rank-merge fold: parity + timings at the bench's shard sizes
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_full_size_oracle.py tests/test_gpu_store.py -x -q -m gpu > gpurun_out/bm25_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/bm25_tests.log
tail -3 gpurun_out/bm25_tests.log
timeout 600 python tools/bm25_sweep.py --docs 6250000 > gpurun_out/bm25_sweep.log 2>&1
timeout 600 python tools/bm25_sweep.py --docs 50000000 >> gpurun_out/bm25_sweep.log 2>&1
timeout 600 python tools/bm25_sweep.py --docs 10000000 --batch 1024 >> gpurun_out/bm25_sweep.log 2>&1
timeout 600 python tools/bm25_sweep.py --docs 1000000 --batch 1 >> gpurun_out/bm25_sweep.log 2>&1
timeout 600 python tools/bm25_sweep.py --docs 1000000 --batch 16 >> gpurun_out/bm25_sweep.log 2>&1
cat gpurun_out/bm25_sweep.log
timeout 300 python tools/bm25_probe.py --docs 10000000
