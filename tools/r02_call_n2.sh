#!/bin/bash
# 2 GPUs: sharded parity (both transports) + full-size oracle at the configs[4] shard size (2 x 12.5M docs) + the headline bench at N=2
set -u
mkdir -p gpurun_out
echo "== multigpu_check x2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/multigpu_check.py > gpurun_out/mg_check2.log 2>&1; echo "exit $?"; grep -E "ok|Error|error|assert" gpurun_out/mg_check2.log | tail -5
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?"; tail -1 gpurun_out/bench_n2.json | cut -c1-200; tail -3 gpurun_out/bench_n2.err
