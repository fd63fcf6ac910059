#!/usr/bin/env python3
"""Run one configuration once (plus one warm-up) for ncu: prof_one.py CONFIG SCALE [TMAX] [rng]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda_pro_cell_b200 import api, synth  # noqa: E402

cfg, scale = int(sys.argv[1]), float(sys.argv[2])
w = synth.workload(cfg, scale)
if len(sys.argv) > 3 and float(sys.argv[3]) > 0:
    w.t_max = float(sys.argv[3])
if len(sys.argv) > 4 and sys.argv[4] == "rng":
    print(api.rng_ceiling(0, 2048))
plan = api.Plan(w.values, w.freqs, w.phi)
eng = api.Engine(0)
eng.load(plan, w.types, w.t_max, w.seed)
for _ in range(2):
    eng.run()
    r = eng.finish(fetch=False)
print(r.stats)
