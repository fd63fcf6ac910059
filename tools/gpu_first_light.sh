#!/bin/bash
# first GPU contact; every step under its own timeout so a hung kernel cannot eat the box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 900 python tools/first_light.py > gpurun_out/first_light.log 2>&1; echo "first_light rc=$?" | tee -a gpurun_out/first_light.log
tail -5 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; tail -30 gpurun_out/first_light.log
