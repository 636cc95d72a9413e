#!/bin/bash
# round 2, GPU call 1: parity suite on the reworked kernels, the overlap sweep, the new headline bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== pytest (without the full-size file)"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size_oracle.py > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest.log
echo "== pytest full-size oracle"; timeout 900 python -m pytest tests/test_gpu_full_size_oracle.py -m gpu -q --durations=10 > gpurun_out/pytest_full.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_full.log
echo "== overlap sweep"; timeout 600 python tools/overlap_sweep.py > gpurun_out/overlap_sweep.json 2> gpurun_out/overlap_sweep.err; echo "sweep exit $?"; cat gpurun_out/overlap_sweep.json; tail -5 gpurun_out/overlap_sweep.err
echo "== bench (headline, 50M docs)"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
ls -la gpurun_out
