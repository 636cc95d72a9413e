#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/graph_probe.py > gpurun_out/graph_probe.log 2>&1; echo "exit $?"; tail -8 gpurun_out/graph_probe.log
