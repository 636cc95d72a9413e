#!/bin/bash
# 4 GPUs: the headline bench at N=4 (12.5M documents per GPU) + configs[4] secondary
set -u
mkdir -p gpurun_out
echo "== bench N=4"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "exit $?"; tail -1 gpurun_out/bench_n4.json | cut -c1-200; tail -3 gpurun_out/bench_n4.err
