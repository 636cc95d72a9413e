#!/bin/bash
# BM25 kernel iteration: parity suites + A/B of build variants on the same box (tools/probes/<name>/libopenintel_gpu.so)
set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/bm25_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/bm25_tests.log
tail -5 gpurun_out/bm25_tests.log
OI_GPU_LIB=$GRAFT_REPO_ROOT/tools/probes/stats/libopenintel_gpu.so timeout 300 python tools/bm25_probe.py --once --batch 1024 > gpurun_out/bm25_stats.log 2>&1
OI_GPU_LIB=$GRAFT_REPO_ROOT/tools/probes/stats/libopenintel_gpu.so timeout 300 python tools/bm25_probe.py --once --batch 256 --docs 6250000 >> gpurun_out/bm25_stats.log 2>&1
grep "bm25 stats" gpurun_out/bm25_stats.log | sort | uniq -c | sort -rn | head -12
for docs in 10000000 6250000; do
  for v in r2a intree; do
    if [ $v = intree ]; then unset OI_GPU_LIB; else export OI_GPU_LIB=$GRAFT_REPO_ROOT/tools/probes/$v/libopenintel_gpu.so; fi
    timeout 300 python tools/bm25_probe.py --docs $docs >> gpurun_out/bm25_ab.log 2>&1
  done
done
unset OI_GPU_LIB
python - <<'PY'
import json
for l in open("gpurun_out/bm25_ab.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-40s %9d b1024 %.3f  uni %.3f  b256 %.3f  uni %.3f" % (d["lib"][-40:], d["n_docs"], min(d["batch1024_zipf"]["ms"]), min(d["batch1024_uniform"]["ms"]), min(d["batch256_zipf"]["ms"]), min(d["batch256_uniform"]["ms"])))
PY
