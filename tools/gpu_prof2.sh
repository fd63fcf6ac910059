#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_c2_b python tools/prof_one.py 2 1.0 > gpurun_out/prof_c2_b.log 2>&1; echo "c2 rc=$?"
tail -n 2 gpurun_out/prof_c2_b.log
