#!/bin/bash
# GPU PostAnalyzer: parity tests + C++ throughput with and without the fused summary; CTA-pair GEMM and graph-capture tests
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lexicon.py tests/test_gpu_host_demo.py tests/test_gpu_full_size_oracle.py -x -q -m gpu > gpurun_out/lex_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/lex_tests.log
for n in 2000 200000 2000000; do timeout 300 openintel_b200/host/host_demo --lexicon-bench $n 5; done 2>&1 | tee gpurun_out/lexicon_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:lexicon_kernel openintel_b200/host/host_demo --lexicon-bench 2000000 2 2>/dev/null | grep lexicon_kernel | tail -2
