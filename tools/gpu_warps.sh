#!/bin/bash
# CTA shape A/B: 16 / 24 / 32 warps per CTA on config 2 (full) and config 4 (1e3 cells, t_max 600)
cd "$(dirname "$0")/.."
for wps in 16 24 32; do
  for rep in 1 2; do
    a=$(PROCELL_COOP_WARPS=$wps python tools/prof_one.py 2 1.0 | tail -1 | grep -o "'kernel_ms': [0-9.]*")
    b=$(PROCELL_COOP_WARPS=$wps python tools/prof_one.py 4 0.1 600 | tail -1 | grep -o "'kernel_ms': [0-9.]*")
    echo "warps=$wps rep$rep cfg2 $a cfg4 $b"
  done
done
