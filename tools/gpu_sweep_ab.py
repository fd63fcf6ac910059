#!/usr/bin/env python3
"""Config 5 (1024 parameter sets on one histogram) with the hashed {key, count} cache against the set-relative direct
table (PROCELL_SWEEP_DIRECT=1): kernel time, a checksum of the count tensor, idle time.  No torch, seconds long.
  python tools/gpu_sweep_ab.py [scale ...]        default: 0.1 1.0"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
from cuda_pro_cell_b200 import api, synth  # noqa: E402

scales = [float(x) for x in sys.argv[1:]] or [0.1, 1.0]
out, t00 = {}, time.time()
for scale in scales:
    w = synth.workload(5, scale)
    plan = api.Plan(w.values, w.freqs, w.phi)
    for mode in ("0", "1"):
        os.environ["PROCELL_SWEEP_DIRECT"] = mode
        eng = api.Engine(0)
        eng.load(plan, w.types, w.t_max, w.seed)
        best = None
        for rep in range(2):
            eng.run()
            r = eng.finish(fetch=(rep == 1))
            best = r.stats if best is None or r.stats["kernel_ms"] < best["kernel_ms"] else best
        flat = r.counts.reshape(-1).astype(np.uint64)
        chk = int((flat * (np.arange(flat.size, dtype=np.uint64) % np.uint64(1000003) + np.uint64(1))).sum() % np.uint64(1 << 61))
        key = "scale%g_direct%s" % (scale, mode)
        out[key] = dict(kernel_ms=best["kernel_ms"], divisions=int(r.divisions.sum()), smem=best["smem_bytes"], checksum=chk,
                        idle_warp_us=best["idle_warp_us"], idle_waits=best["idle_waits"], donations=best["donations"])
        print(key, out[key], "t=%.1fs" % (time.time() - t00), flush=True)
        eng.close()
    a, b = out["scale%g_direct0" % scale], out["scale%g_direct1" % scale]
    print("scale %g: same tensor: %s, direct / hashed kernel time = %.3f" % (scale, a["checksum"] == b["checksum"] and a["divisions"] == b["divisions"], b["kernel_ms"] / a["kernel_ms"]), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "config5_setdirect_ab.json", "w"), indent=1)
