#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-p}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rng_ceiling -s 1 -c 1 -f -o gpurun_out/prof_rng_$TAG python tools/prof_one.py 1 1.0 0 rng > gpurun_out/prof_rng_$TAG.log 2>&1; echo "rng rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_c2_$TAG python tools/prof_one.py 2 1.0 > gpurun_out/prof_c2_$TAG.log 2>&1; echo "c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_c4_$TAG python tools/prof_one.py 4 0.1 600 > gpurun_out/prof_c4_$TAG.log 2>&1; echo "c4 rc=$?"
for f in rng c2 c4; do tail -n 2 gpurun_out/prof_${f}_$TAG.log; done
