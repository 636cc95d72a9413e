#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest all gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_all.log
echo "== lexicon bench (C++ host_demo)"; for n in 2000 200000 2000000; do timeout 120 openintel_b200/host/host_demo --lexicon-bench $n 10 2>&1 | tee -a gpurun_out/lexicon_bench.log; done
echo "== lexicon ncu"
C="openintel_b200/host/host_demo --lexicon-bench 2000000 2"
timeout 120 $C > gpurun_out/plain_lx.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lexicon_kernel -s 1 -c 1 -f -o gpurun_out/prof_lexicon $C > gpurun_out/ncu_lx.log 2>&1
echo "ncu rc $?"
echo "== hybrid probe 6.25M"; timeout 300 python tools/hybrid_probe.py 2>gpurun_out/hp.err | tee gpurun_out/hybrid_probe_6m.json
