#!/bin/bash
# quick GPU visit: parity tests + timing survey
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/first_light.py 2>&1 | tee gpurun_out/first_light.log | grep -v "^ref run" | tail -30
