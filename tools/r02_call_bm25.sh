#!/bin/bash
# BM25 after the threshold fold: parity, headline bench, ncu capture of the kernel at configs[2]
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/bm25_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/bm25_tests.log
tail -3 gpurun_out/bm25_tests.log
timeout 900 python bench.py > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_n1_b.json
C="python tools/bm25_probe.py --once --batch 1024"
timeout 300 $C > gpurun_out/plain_b.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:bm25_blocked -s 2 -c 1 -o gpurun_out/r02b_prof_bm25 $C > gpurun_out/ncu_b.log 2>&1
echo "ncu rc=$?"
