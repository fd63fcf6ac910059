#!/usr/bin/env python3
"""First GPU contact: timings of the kernels on the BASELINE configs and the RNG ceiling.  Writes
gpurun_out/first_light.json.  (The reference binary is timed by `bench.py --impl reference` only.)"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda_pro_cell_b200 import api, synth  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
res = {}


def timed(plan, w, kernel, reps=3):
    eng = api.Engine(0)
    eng.load(plan, w.types, w.t_max, w.seed, kernel=kernel)
    best = None
    for _ in range(reps):
        eng.run()
        r = eng.finish(fetch=False)
        best = r.stats if best is None or r.stats["kernel_ms"] < best["kernel_ms"] else best
    eng.close()
    return best


import os
for cfg, scale in ((1, 1.0), (2, 1.0), (3, 0.1), (4, 0.1), (5, 0.02)):
    w = synth.workload(cfg, scale)
    if cfg == 4:
        w.t_max = 600.0
    plan = api.Plan(w.values, w.freqs, w.phi)
    for kname, k in (("coop32", 0),):
        os.environ["PROCELL_COOP_WARPS"] = kname[4:6]
        os.environ["PROCELL_NO_DONATE"] = "1" if kname.endswith("nodonate") else "0"
        try:
            st = timed(plan, w, k)
            st["Gdiv_per_s"] = st["divisions"] / st["kernel_ms"] / 1e6
            res["cfg%d_%s" % (cfg, kname)] = st
            print(cfg, kname, st, flush=True)
        except Exception as e:  # keep going: this is a survey run
            res["cfg%d_%s" % (cfg, kname)] = {"error": str(e)}
            print(cfg, kname, "ERROR", e, flush=True)

try:
    ms, pairs = api.rng_ceiling(0, 4096)
    res["rng_ceiling"] = {"ms": ms, "pairs": pairs, "Gpairs_per_s": pairs / ms / 1e6}
    print("rng ceiling", res["rng_ceiling"], flush=True)
except Exception as e:
    res["rng_ceiling"] = {"error": str(e)}

(OUT / "first_light.json").write_text(json.dumps(res, indent=1))
print("done")
