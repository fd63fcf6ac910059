#!/bin/bash
# last GPU visit of a round: parity suite on the committed build, then compute-sanitizer memcheck on small runs that
# cover direct + hashed histograms, spill + donation, time series and the CTA-shape variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
timeout 100 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/pytest_gpu_last.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-T0 ))s"
tail -3 gpurun_out/pytest_gpu_last.log
cat > /tmp/san.py <<'PY'
import os, sys; sys.path.insert(0,'.')
import numpy as np
from cuda_pro_cell_b200 import api, synth
v,f=synth.synthetic_histogram(2000)
plan=api.Plan(v,f,0.5)
r=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3)                 # direct histogram, PLAIN instance
r2=api.proliferate(plan,synth.sweep_types(1024)[::256],80.0,4)        # hashed histogram + batches
r3=api.proliferate(api.Plan(np.array([1000.0]),np.array([3],dtype=np.uint64),1e-6),[[(1.0,24.0,4.0)]],250.0,5)  # spill + donation
r4=api.proliferate(plan,[synth.TYPES_CONFIG2],80.0,3,checkpoints=[20.0,80.0])
os.environ["PROCELL_COOP_NPL"]="2"
r5=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3)
print("ok",int(r.divisions.sum()),int(r2.divisions.sum()),int(r3.divisions.sum()),r3.stats['donations'],int(r4.counts.sum()),bool(np.array_equal(r.counts,r5.counts)))
PY
timeout 70 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_memcheck_last.log 2>&1; echo "memcheck rc=$? t=$(( $(date +%s)-T0 ))s"
grep -E "ERROR SUMMARY|^ok|Invalid|out of bounds" gpurun_out/sanitizer_memcheck_last.log | head -6
