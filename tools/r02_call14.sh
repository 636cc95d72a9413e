#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== hybrid probe 6.25M"; timeout 300 python tools/hybrid_probe.py 2>gpurun_out/hp.err | tee gpurun_out/hybrid_probe_6m.json
echo "== launch list"
timeout 300 python tools/hybrid_probe.py --once > gpurun_out/plain_h.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_hybrid_6m.csv python tools/hybrid_probe.py --once > gpurun_out/ncu_h.log 2>&1
echo "ncu exit $?"; grep -v "^==" gpurun_out/launches_hybrid_6m.csv | awk -F'","' '{print substr($5,1,70), $(NF)}' | tail -16
