#!/usr/bin/env python3
"""Seconds-long parity check of the kernel modes that were written without GPU access (no torch, no pytest):
  python tests/gpu_quick_new_modes.py subtree | setdirect
Prints one PASS / FAIL line per case.  Run each mode in its own process under `timeout`, with PROCELL_WATCHDOG_S small."""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))      # the oracle binding lives with the tests: this script is a test tool
import numpy as np  # noqa: E402
from cuda_pro_cell_b200 import api, synth  # noqa: E402
import oracle_lib  # noqa: E402  (the checker; this script is a test tool)

mode = sys.argv[1]
t00 = time.time()
if mode == "subtree":
    v, f = synth.synthetic_histogram(800)
    types = np.array([synth.TYPES_CONFIG4])
    for phi in (1e-3, 1e-7):
        plan, oplan = api.Plan(v, f, phi), oracle_lib.OraclePlan(v, f, phi)
        whole = oracle_lib.simulate(oplan, types, 360.0, 4)
        for world, level in ((3, 4), (8, 6), (2, 1)):
            total, ok = np.zeros_like(whole["counts"]), True
            for r in range(world):
                got = api.proliferate(plan, types, 360.0, 4, shard=(r, world, 32), shard_level=level)
                want = oracle_lib.simulate(oplan, types, 360.0, 4, shard=(r, world, 32), shard_level=level)
                ok = ok and np.array_equal(got.counts, want["counts"]) and int(got.divisions[0]) == int(want["divisions"][0])
                total += got.counts
            ok = ok and np.array_equal(total, whole["counts"])
            print("%s subtree phi=%g world=%d level=%d divisions=%d smem=%d t=%.1fs" % ("PASS" if ok else "FAIL", phi, world, level, int(whole["divisions"][0]), got.stats["smem_bytes"], time.time() - t00), flush=True)
elif mode == "setdirect":
    cases = [("sweep16", 1500, synth.sweep_types(1024)[::64], 168.0, 0.5, (0, 1, 0)),
             ("many_small", 300, synth.sweep_types(1024)[::8], 168.0, 0.5, (0, 1, 0)),
             ("two_deep", 600, np.array([synth.TYPES_CONFIG4, [(0.02, 20.0, 3.0), (0.28, 86.3, 26.8), (0.70, -1.0, -1.0)]]), 330.0, 1e-7, (0, 1, 0)),
             ("sharded", 3000, synth.sweep_types(1024)[::128], 200.0, 0.5, (1, 3, 32))]
    for name, n, types, t_max, phi, shard in cases:
        v, f = synth.synthetic_histogram(n)
        plan, oplan = api.Plan(v, f, phi), oracle_lib.OraclePlan(v, f, phi)
        want = oracle_lib.simulate(oplan, types, t_max, 5, shard=shard if shard[1] > 1 else (0, 1, 1))
        os.environ["PROCELL_SWEEP_DIRECT"] = "0"
        h = api.proliferate(plan, types, t_max, 5, shard=shard)
        os.environ["PROCELL_SWEEP_DIRECT"] = "1"
        g = api.proliferate(plan, types, t_max, 5, shard=shard)
        ok = np.array_equal(g.counts, want["counts"]) and np.array_equal(g.divisions, want["divisions"])
        sel = g.stats["smem_bytes"] != h.stats["smem_bytes"]
        print("%s setdirect %s sets=%d divisions=%d selected=%s hashed_ok=%s ms hashed/direct=%.3f/%.3f t=%.1fs" % (
            "PASS" if ok and sel else "FAIL", name, len(types), int(want["divisions"].sum()), sel,
            np.array_equal(h.counts, want["counts"]), h.stats["kernel_ms"], g.stats["kernel_ms"], time.time() - t00), flush=True)
print("done", mode, "%.1fs" % (time.time() - t00))
