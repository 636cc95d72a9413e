#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bm25 tests"; timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_full_size_oracle.py -m gpu -x -q > gpurun_out/pytest9.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest9.log
echo "== bm25 probe"; timeout 600 python tools/bm25_probe.py 2>gpurun_out/bm25_probe.err | tee gpurun_out/bm25_probe.json; tail -3 gpurun_out/bm25_probe.err
