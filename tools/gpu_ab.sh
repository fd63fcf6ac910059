#!/bin/bash
# A/B timing of library variants on config 2, config 4 (t_max 600, 1e3 cells) and config 5 at 1/10 size (kernel_ms of the second run of each process)
cd "$(dirname "$0")/.."
for lib in "$@"; do
  for rep in 1 2 3; do
    a=$(PROCELL_LIB=$lib python tools/prof_one.py 2 1.0 | tail -1 | grep -o "'kernel_ms': [0-9.]*")
    b=$(PROCELL_LIB=$lib python tools/prof_one.py 4 0.1 600 | tail -1 | grep -o "'kernel_ms': [0-9.]*")
    c=$(PROCELL_LIB=$lib python tools/prof_one.py 5 0.1 | tail -1 | grep -o "'kernel_ms': [0-9.]*")
    d=$(PROCELL_LIB=$lib python tools/prof_one.py 3 0.1 | tail -1 | grep -o "'kernel_ms': [0-9.]*")
    echo "$lib rep$rep cfg2 $a cfg4 $b cfg5/10 $c cfg3/10 $d"
  done
done
