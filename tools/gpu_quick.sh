#!/bin/bash
# quick GPU visit: parity tests + timing survey
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/first_light.py 2>&1 | tee gpurun_out/first_light.log | grep -v "^ref run" | tail -30
