#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python tests/golden/make_ref_fixtures.py 2>&1 | tail -30
timeout 600 python - <<'PY'
import sys; sys.path.insert(0,'.')
from cuda_pro_cell_b200 import api, synth
import numpy as np, time
w=synth.workload(5,1.0)
plan=api.Plan(w.values,w.freqs,w.phi)
eng=api.Engine(0); eng.load(plan,w.types,w.t_max,w.seed)
for i in range(2):
    eng.run(); r=eng.finish(fetch=False); print("config5 full", r.stats, r.stats['divisions']/r.stats['kernel_ms']/1e6, "Gdiv/s", flush=True)
PY
