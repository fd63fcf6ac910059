#!/usr/bin/env python3
"""Config 5 at full size with a target histogram: what the sweep fitness costs in the launch and as a separate pass.
  gpurun -- python tools/fitness_timing.py   -> gpurun_out/fitness_timing.json"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
from cuda_pro_cell_b200 import api, synth  # noqa: E402

w = synth.workload(5)
plan = api.Plan(w.values, w.freqs, w.phi)
tvalues = plan.row_value[3::4].copy()
tfreqs = np.arange(1, len(tvalues) + 1, dtype=np.uint64)
out = {}
for label, env in (("no_target", None), ("fitness_in_launch", "1"), ("separate_pass", "0")):
    if env is not None:
        os.environ["PROCELL_FITNESS_FUSED"] = env
    eng = api.Engine(0)
    eng.load(plan, w.types, w.t_max, w.seed)
    if env is not None:
        eng.set_target(tvalues, tfreqs)
    best, fit_ms, fit = None, [], None
    for _ in range(3):
        eng.run()
        r = eng.finish(fetch=False)
        best = r.stats["kernel_ms"] if best is None else min(best, r.stats["kernel_ms"])
        if env is not None:
            t0 = time.perf_counter()
            fit = eng.fitness()
            fit_ms.append(1e3 * (time.perf_counter() - t0))
    out[label] = {"kernel_ms": best, "fitness_call_ms": min(fit_ms) if fit_ms else None,
                  "in_launch": eng.fitness_in_launch() if env is not None else None,
                  "fitness_checksum": None if fit is None else float(np.sum(fit * np.arange(1, len(fit) + 1)))}
    eng.close()
    print(label, out[label], flush=True)
out["bytes_leaving_the_gpu"] = {"fitness": int(w.types.shape[0]) * 8, "count_tensor": int(w.types.shape[0]) * plan.n_keys * w.types.shape[1] * 8}
json.dump(out, open(ROOT / "gpurun_out" / "fitness_timing.json", "w"), indent=1)
