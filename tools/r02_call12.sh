#!/bin/bash
# 2 GPUs: sharded parity check + the headline bench at N=2 (with the configs[4] secondary) + reference arm under torchrun
set -u
mkdir -p gpurun_out
echo "== multigpu_check"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/mg_check2.log 2>&1; echo "exit $?"; tail -3 gpurun_out/mg_check2.log
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?"; tail -1 gpurun_out/bench_n2.json | cut -c1-3000; tail -4 gpurun_out/bench_n2.err
echo "== reference arm N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; echo "exit $?"; tail -1 gpurun_out/bench_ref2.json | cut -c1-600
