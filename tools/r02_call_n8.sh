#!/bin/bash
# 8 GPUs: sharded parity (both transports) + configs[4] full-size oracle (8 x 12.5M = 100M docs) + the headline bench at N=8
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc8.txt
echo "== multigpu_check x8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/multigpu_check.py > gpurun_out/mg_check8.log 2>&1; echo "exit $?"; grep -E "ok|Error|error|assert" gpurun_out/mg_check8.log | tail -5
echo "== bench N=8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "exit $?"; tail -1 gpurun_out/bench_n8.json | cut -c1-200; tail -3 gpurun_out/bench_n8.err
