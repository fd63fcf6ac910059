#!/bin/bash
# A/B timing of library variants on config 2 and config 4 (kernel_ms of the second run of each process)
cd "$(dirname "$0")/.."
for lib in "$@"; do
  for rep in 1 2 3; do
    a=$(PROCELL_LIB=$lib python tools/prof_one.py 2 1.0 | tail -1 | grep -o "'kernel_ms': [0-9.]*")
    b=$(PROCELL_LIB=$lib python tools/prof_one.py 4 0.1 600 | tail -1 | grep -o "'kernel_ms': [0-9.]*")


    echo "$lib rep$rep cfg2 $a cfg4 $b"
  done
done
