#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
echo "== pytest all gpu"; timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?"; tail -16 gpurun_out/pytest_all.log
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
