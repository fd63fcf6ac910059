#!/usr/bin/env python3
"""A/B of the kernel's run-time shape knobs in ONE process (one CUDA context): for every knob setting and workload,
load the engine (the knobs are read at load time), run REPS times, report min / median kernel_ms and a checksum of
the count tensor (all settings must agree: results do not depend on scheduling).

  python tools/ab_knobs.py [REPS] [default,w16,w24]     # prints one JSON line per (knob, workload)
  PROCELL_LIB=libprocell_b200_x.so python tools/ab_knobs.py 5 default       # another in-tree build of the library
"""
import json
import os
import statistics
import sys
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda_pro_cell_b200 import api, synth  # noqa: E402

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ALL_KNOBS = {"default": {}, "w16": {"PROCELL_COOP_WARPS": "16"},
             "w24": {"PROCELL_COOP_WARPS": "24"},
             "nomerge": {"PROCELL_LEAF_MERGE": "0"}, "merge": {"PROCELL_LEAF_MERGE": "1"}}   # kernel MODE 3 forced off / on
KNOBS = [ALL_KNOBS[k] for k in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("default", "w16", "w24"))]
WORK = [(2, 1.0, 0.0), (2, 0.1, 0.0), (3, 1.0, 0.0), (5, 1.0, 0.0), (4, 0.1, 600.0), (4, 1.0, 0.0)]
if os.environ.get("AB_WORK"):        # e.g. AB_WORK="2:1.0:0,4:0.1:600"
    WORK = [tuple(float(x) if i else int(x) for i, x in enumerate(item.split(":"))) for item in os.environ["AB_WORK"].split(",")]

for cfg, scale, t_override in WORK:
    w = synth.workload(cfg, scale)
    if t_override > 0:
        w.t_max = t_override
    plan = api.Plan(w.values, w.freqs, w.phi)
    for knob in KNOBS:
        for k in ("PROCELL_COOP_NPL", "PROCELL_COOP_WARPS", "PROCELL_LEAF_MERGE"):
            os.environ.pop(k, None)
        os.environ.update(knob)
        eng = api.Engine(0)
        eng.load(plan, w.types, w.t_max, w.seed)
        ms, crc, div = [], None, None
        for i in range(REPS + 1):
            eng.run()
            r = eng.finish(fetch=(i == 0))
            if i == 0:
                crc = zlib.crc32(r.counts.tobytes())
                div = int(r.divisions.sum())
            else:
                ms.append(r.stats["kernel_ms"])
        print(json.dumps({"lib": os.environ.get("PROCELL_LIB", "libprocell_b200.so"), "config": cfg, "scale": scale, "t_max": w.t_max, "knob": knob, "divisions": div, "crc": crc,
                          "ms_min": min(ms), "ms_med": statistics.median(ms), "block": r.stats["block"],
                          "seed_phase_us": r.stats["seed_phase_us"], "span_us": r.stats["total_us"],
                          "idle_us_per_warp": r.stats["idle_warp_us"] / max(1, r.stats["grid"] * r.stats["block"] // 32),
                          "donations": r.stats["donations"],
                          "Gdiv_s": div / min(ms) / 1e6}), flush=True)
        eng.close()
