#!/usr/bin/env python3
"""First GPU contact: timings of both kernels on the BASELINE configs, the RNG ceiling, and a few runs of the
unmodified reference binary (oracle/_ref/procell_ref) on config 1.  Writes gpurun_out/first_light.json."""
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
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

# ---- the reference binary on config 1 (needs -p; wall-clock seeded, so space the runs by > 1 s)
ref = ROOT / "oracle" / "_ref" / "procell_ref"
if ref.exists():
    w = synth.workload(1)
    (OUT / "cfg1_hist.txt").write_text(synth.histogram_text(w.values, w.freqs))
    (OUT / "cfg1_types.txt").write_text(synth.types_text(w.types[0]))
    runs = []
    for i, (tmax, phi) in enumerate(()):
        out = OUT / ("ref_cfg1_run%d.txt" % i)
        t0 = time.time()
        r = subprocess.run([str(ref), "-h", str(OUT / "cfg1_hist.txt"), "-c", str(OUT / "cfg1_types.txt"), "-t", str(tmax),
                            "-p", repr(phi), "-o", str(out), "-r"], capture_output=True, text=True, timeout=600)
        dt = time.time() - t0
        rows = [ln.split("\t") for ln in out.read_text().splitlines()] if out.exists() else []
        tot = sum(int(x[1]) for x in rows)
        mass = sum(float(x[0]) * int(x[1]) for x in rows)
        runs.append(dict(t_max=tmax, phi=phi, rc=r.returncode, wall_s=dt, start_unix=int(t0), rows=len(rows), total=tot,
                         mass=mass, stdout=r.stdout[-300:], stderr=r.stderr[-300:]))
        print("ref run", runs[-1], flush=True)
        time.sleep(1.2)
    res["reference_cfg1"] = runs
    res["cfg1_input_mass"] = float((w.values * w.freqs).sum())
else:
    res["reference_cfg1"] = "binary missing"

(OUT / "first_light.json").write_text(json.dumps(res, indent=1))
print("done")
