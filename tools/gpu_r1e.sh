#!/bin/bash
# one short GPU visit (13 GPU-minutes were left): (A) the whole parity suite on the default build after the
# divide_iteration refactor and the new text I/O, (B) A/B of the kernel shapes incl. the two-nodes-per-lane instance,
# (C) the parity suite once more with that instance selected
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q --timeout 90 --durations=6 > gpurun_out/pytest_gpu_r1e.log 2>&1; echo "A pytest rc=$? t=$(( $(date +%s)-T0 ))s"
tail -12 gpurun_out/pytest_gpu_r1e.log
timeout 150 python tools/ab_knobs.py 5 > gpurun_out/ab_knobs_r1e.jsonl 2> gpurun_out/ab_knobs_r1e.err; echo "B ab rc=$? t=$(( $(date +%s)-T0 ))s"
cat gpurun_out/ab_knobs_r1e.jsonl; tail -3 gpurun_out/ab_knobs_r1e.err
PROCELL_COOP_NPL=2 timeout 240 python -m pytest tests -m gpu -x -q --timeout 90 -k "not reference and not cli and not multi_gpu" > gpurun_out/pytest_gpu_npl2_r1e.log 2>&1; echo "C pytest npl2 rc=$? t=$(( $(date +%s)-T0 ))s"
tail -6 gpurun_out/pytest_gpu_npl2_r1e.log
