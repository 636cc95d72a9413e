#!/bin/bash
# final single-GPU validation of the round: full GPU suite, smoke, bench (both arms)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench exit $?"; tail -1 gpurun_out/final_bench_n1.json | cut -c1-400
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref exit $?"; tail -1 gpurun_out/final_bench_ref.json | cut -c1-600
