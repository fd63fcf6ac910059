#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck) on small runs that cover every kernel mode:
#   gpurun --timeout 600 -- 'bash tools/gpu_sanitize.sh TAG'      -> gpurun_out/sanitizer_{memcheck,racecheck,initcheck}_TAG.log
# direct table in slot layout (PLAIN), hashed cache + batches, spill + donation, time series, the 16-warp shape,
# subtree sharding (MODE 1), set-relative sweep table with rendezvous (MODE 2), sweep fitness in the launch,
# merged leaf counts (MODE 3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2}
cat > /tmp/san.py <<'PY'
import os, sys; sys.path.insert(0,'.')
import numpy as np
from cuda_pro_cell_b200 import api, synth
v,f=synth.synthetic_histogram(2000)
plan=api.Plan(v,f,0.5)
r=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3)                 # direct table in slot layout, PLAIN instance
r2=api.proliferate(plan,synth.sweep_types(1024)[::256],80.0,4)        # hashed cache + batches
r3=api.proliferate(api.Plan(np.array([1000.0]),np.array([3],dtype=np.uint64),1e-6),[[(1.0,24.0,4.0)]],250.0,5)  # spill + donation
r4=api.proliferate(plan,[synth.TYPES_CONFIG2],80.0,3,checkpoints=[20.0,80.0])
r6=api.proliferate(plan,[synth.TYPES_CONFIG4],200.0,3,shard=(1,3,32),shard_level=4)    # subtree sharding (MODE 1)
os.environ["PROCELL_SWEEP_DIRECT"]="1"
eng=api.Engine(0); types=synth.sweep_types(1024)[::128]
eng.load(plan,types,80.0,4); eng.set_target(plan.row_value[::3].copy(), np.arange(1,len(plan.row_value[::3])+1,dtype=np.uint64))
eng.run(); r7=eng.finish(); fit=eng.fitness(); fused=eng.fitness_in_launch(); eng.close()   # MODE 2 + fitness in the launch
del os.environ["PROCELL_SWEEP_DIRECT"]
os.environ["PROCELL_LEAF_MERGE"]="1"
r8=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3)                 # merged leaf counts (MODE 3), forced on this small input
del os.environ["PROCELL_LEAF_MERGE"]
assert np.array_equal(r8.counts,r.counts)
os.environ["PROCELL_COOP_WARPS"]="16"
r5=api.proliferate(plan,[synth.TYPES_CONFIG2],100.0,3)
print("ok",int(r.divisions.sum()),int(r2.divisions.sum()),int(r3.divisions.sum()),r3.stats['donations'],int(r4.counts.sum()),
      int(r6.divisions.sum()),int(r7.divisions.sum()),fused,float(fit.min()),bool(np.array_equal(r.counts,r5.counts)))
PY
for tool in memcheck racecheck initcheck; do
  timeout 150 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok" gpurun_out/sanitizer_${tool}_$TAG.log | head -4
done
grep -E "hazard" gpurun_out/sanitizer_racecheck_$TAG.log | sed 's/0x[0-9a-f]*//g' | sed 's/procell_b200::<unnamed>:://g' | cut -c1-220 | sort | uniq -c | sort -rn | head -24
