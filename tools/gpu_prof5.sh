#!/bin/bash
# ncu --set full of the sweep workload (config 5 at 1/10 size: 1024 sets x 1e5 cells) and of config 3 at 1/10 size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-p}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_c5_$TAG python tools/prof_one.py 5 0.1 > gpurun_out/prof_c5_$TAG.log 2>&1; echo "c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_c3_$TAG python tools/prof_one.py 3 0.1 > gpurun_out/prof_c3_$TAG.log 2>&1; echo "c3 rc=$?"
for f in c5 c3; do tail -n 1 gpurun_out/prof_${f}_$TAG.log; done
