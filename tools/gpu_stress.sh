#!/bin/bash
# repeat the GPU parity suite to flush out rare races; every test has its own timeout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PROCELL_WATCHDOG_S=20
for i in $(seq 1 ${1:-6}); do
  timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/stress_$i.log 2>&1; echo "round $i rc=$? $(tail -1 gpurun_out/stress_$i.log | cut -c1-200)"
done
