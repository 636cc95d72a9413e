#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest lexicon + host demo"; timeout 600 python -m pytest tests/test_gpu_lexicon.py tests/test_gpu_host_demo.py -m gpu -x -q > gpurun_out/pytest_lx.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_lx.log
echo "== lexicon bench (C++ host_demo)"; rm -f gpurun_out/lexicon_bench.log; for n in 2000 200000 2000000; do timeout 120 openintel_b200/host/host_demo --lexicon-bench $n 10 2>&1 | tee -a gpurun_out/lexicon_bench.log; done
echo "== lexicon ncu"
C="openintel_b200/host/host_demo --lexicon-bench 2000000 2"
timeout 120 $C > gpurun_out/plain_lx.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lexicon_kernel -s 1 -c 1 -f -o gpurun_out/prof_lexicon $C > gpurun_out/ncu_lx.log 2>&1
echo "ncu rc $?"
