#!/bin/bash
# one GPU visit: parity tests, bench (ours + reference arm), ncu launch list and a full capture of the top kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
tail -4 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
tail -5 gpurun_out/ncu_full_$TAG.log
