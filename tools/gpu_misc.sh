#!/bin/bash
# (1) config 5 full with 24 vs 32 warps per CTA, (2) compute-sanitizer memcheck + initcheck on small runs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 24 32; do
  PROCELL_COOP_WARPS=$w timeout 200 python tools/prof_one.py 5 1.0 | tail -1 | grep -o "'kernel_ms': [0-9.]*" | sed "s/^/cfg5 warps=$w /"
done
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from cuda_pro_cell_b200 import api, synth
v,f=synth.synthetic_histogram(3000)
plan=api.Plan(v,f,0.5)
r=api.proliferate(plan,[synth.TYPES_CONFIG2],120.0,3)                 # direct histogram
r2=api.proliferate(plan,synth.sweep_types(1024)[::128],100.0,4)       # hashed histogram + batches
r3=api.proliferate(api.Plan(np.array([1000.0]),np.array([4],dtype=np.uint64),1e-6),[[(1.0,24.0,4.0)]],300.0,5)  # spill + donation
r4=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3,checkpoints=[20.0,100.0])
r5=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3,kernel=1)
print("ok",int(r.divisions.sum()),int(r2.divisions.sum()),int(r3.divisions.sum()),r3.stats['donations'],int(r4.counts.sum()),int(r5.divisions.sum()))
PY
for tool in memcheck initcheck; do
  timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|^ok|Invalid|Uninitialized" gpurun_out/sanitizer_$tool.log | head -5
done
