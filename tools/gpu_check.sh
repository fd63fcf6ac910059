#!/bin/bash
# parity tests + one short bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/check_pytest.log
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/check_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/check_bench.json').read().strip().splitlines()[-1]); print("value %.4g e2e %.4g ms/step %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["frac"]))
except Exception as e: print("no bench line", e)
PY
