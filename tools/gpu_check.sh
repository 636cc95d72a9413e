#!/bin/bash
# tools/gpu_check.sh — one gpurun call: smoke, GPU parity tests, bench (both scan variants),
# ncu launch list + one full capture of the scan kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest.log
for v in 0 1; do
  echo "== bench variant $v"
  timeout 600 python bench.py --variant $v --steps 30 --warmup 5 --no-secondary > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench exit $?"
  cat gpurun_out/bench_v$v.json; tail -3 gpurun_out/bench_v$v.err
done
if [ "${SKIP_NCU:-0}" != "1" ]; then
  V=${NCU_VARIANT:-1}; SKIP=${NCU_SKIP:-3}
  echo "== ncu (variant $V)"
  timeout 300 python bench.py --variant $V --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --variant $V --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 300 python bench.py --variant $V --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:cosine_scan -s $SKIP -c 1 -f -o gpurun_out/prof_scan_v$V \
      python bench.py --variant $V --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
ls -la gpurun_out
