#!/bin/bash
# short GPU visit: parity suite on the default build, then A/B of in-tree library builds (default CTA shape unless
# KNOBS is given):  bash tools/gpu_ab2.sh TAG libA.so libB.so ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; shift
T0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-T0 ))s"
tail -4 gpurun_out/pytest_gpu_$TAG.log
: > gpurun_out/ab_$TAG.jsonl
for rep in 1 2; do
  for lib in "$@"; do
    PROCELL_LIB=$lib timeout 120 python tools/ab_knobs.py 7 ${KNOBS:-default} >> gpurun_out/ab_$TAG.jsonl 2>> gpurun_out/ab_$TAG.err
  done
done
echo "ab t=$(( $(date +%s)-T0 ))s"
python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/ab_$TAG.jsonl")]
best={}
for r in rows:
    k=(r["config"],r["scale"],r["lib"],json.dumps(r["knob"]))
    best[k]=min(best.get(k,1e9),r["ms_min"])
crc={}
for r in rows: crc.setdefault((r["config"],r["scale"],r["t_max"]),set()).add((r["crc"],r["divisions"]))
for k in sorted(best): print(k, "%.4f ms"%best[k])
print("crc consistent:", all(len(v)==1 for v in crc.values()))
PY
tail -3 gpurun_out/ab_$TAG.err
