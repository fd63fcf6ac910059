#!/bin/bash
# final checks of a round: smoke(), compute-sanitizer memcheck / racecheck / initcheck on small runs that cover
# direct + hashed histograms, spill + donation, time series, the simple kernel and the one-call entry
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from cuda_pro_cell_b200 import api, synth
v,f=synth.synthetic_histogram(3000)
plan=api.Plan(v,f,0.5)
r=api.proliferate(plan,[synth.TYPES_CONFIG2],120.0,3)                 # direct histogram, PLAIN instance
r2=api.proliferate(plan,synth.sweep_types(1024)[::128],100.0,4)       # hashed histogram + batches
r3=api.proliferate(api.Plan(np.array([1000.0]),np.array([4],dtype=np.uint64),1e-6),[[(1.0,24.0,4.0)]],300.0,5)  # spill + donation
r4=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3,checkpoints=[20.0,100.0])
r5=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3,kernel=1)
r6=api.simulate(v,f,synth.TYPES_CONFIG2,60.0,0.5,7)
print("ok",int(r.divisions.sum()),int(r2.divisions.sum()),int(r3.divisions.sum()),r3.stats['donations'],int(r4.counts.sum()),int(r5.divisions.sum()),r6[3])
PY
for tool in memcheck racecheck initcheck; do
  timeout 500 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|Invalid|Uninitialized|hazard" gpurun_out/sanitizer_$tool.log | head -6
done
