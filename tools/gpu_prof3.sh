#!/bin/bash
# ncu --set full of one launch of the product kernel on configs 2 (full), 3 (1/10) and 5 (1/10)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-p}
T0=$(date +%s)
for spec in "c2 2 1.0" "c3 3 0.1" "c5 5 0.1"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_$1_$TAG python tools/prof_one.py $2 $3 > gpurun_out/prof_$1_$TAG.log 2>&1; echo "$1 rc=$? t=$(( $(date +%s)-T0 ))s"
  tail -n 1 gpurun_out/prof_$1_$TAG.log
done
ls -la gpurun_out/*.ncu-rep
