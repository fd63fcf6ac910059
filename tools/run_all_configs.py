#!/usr/bin/env python3
"""All five BASELINE.json configurations at FULL size on one B200: timing, division counts and size-independent
properties (run-to-run identity, identity across CTA shapes = scheduling independence, -r columns sum to totals,
count/mass conservation where phi never binds).  Writes gpurun_out/all_configs.json."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda_pro_cell_b200 import api, synth  # noqa: E402

out = {}
for cfg in (1, 2, 3, 4, 5):
    w = synth.workload(cfg, 1.0)
    plan = api.Plan(w.values, w.freqs, w.phi)
    res = {}
    runs = []
    for warps in ("32", "32", "16"):
        os.environ["PROCELL_COOP_WARPS"] = warps
        eng = api.Engine(0)
        eng.load(plan, w.types, w.t_max, w.seed)
        eng.run()
        t0 = time.perf_counter()
        r = eng.finish(fetch=True)
        runs.append(r)
        eng.close()
    a, b, c = runs
    res["n_cells"] = int(plan.n_cells)
    res["n_sets"] = int(w.types.shape[0])
    res["n_keys"] = int(plan.n_keys)
    res["phi"] = plan.phi
    res["t_max"] = w.t_max
    res["divisions"] = int(a.divisions.sum())
    res["leaves"] = int(a.counts.sum())
    res["kernel_ms"] = [x.stats["kernel_ms"] for x in runs]
    res["Gdiv_per_s"] = res["divisions"] / min(res["kernel_ms"][:2]) / 1e6
    res["donations"] = [x.stats["donations"] for x in runs]
    res["identical_run_to_run"] = bool(np.array_equal(a.counts, b.counts) and np.array_equal(a.divisions, b.divisions))
    res["identical_16_vs_32_warps"] = bool(np.array_equal(a.counts, c.counts) and np.array_equal(a.divisions, c.divisions))
    res["max_count"] = int(a.counts.max())
    res["exceeds_int32"] = bool(a.counts.max() > 2**31 - 1)
    rf, rr = plan.merge_rows(a.counts[0])
    res["ratio_columns_sum_to_total"] = bool(np.array_equal(rr.sum(axis=1), rf))
    res["leaves_minus_seeds_minus_divisions_set0"] = int(rf.sum()) - int(plan.n_cells) - int(a.divisions[0])
    mass_in = float((w.values * w.freqs).sum())
    res["mass_out_over_in_set0"] = float((rf * plan.row_value).sum()) / mass_in
    if cfg == 1:     # start-up bound: also the per-run time of 100 back-to-back in-process runs
        os.environ["PROCELL_COOP_WARPS"] = "32"
        eng = api.Engine(0)
        eng.load(plan, w.types, w.t_max, w.seed)
        eng.run(); eng.finish(fetch=False)
        t0 = time.perf_counter()
        for i in range(100):
            eng.run(w.seed + i)
        eng.finish(fetch=False)
        res["in_process_loop_us_per_run"] = (time.perf_counter() - t0) * 1e4
        res["in_process_loop_Gdiv_per_s"] = res["divisions"] / (res["in_process_loop_us_per_run"] * 1e-6) / 1e9
        eng.close()
    if cfg in (1, 2, 3):     # end-to-end wall time of the `procell` command line (process start, CUDA context, text I/O)
        import subprocess
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            Path(d, "h.txt").write_text(synth.histogram_text(w.values, w.freqs))
            Path(d, "c.txt").write_text(synth.types_text(w.types[0]))
            cmd = [str(ROOT / "cuda_pro_cell_b200" / "procell"), "-h", d + "/h.txt", "-c", d + "/c.txt", "-t", repr(w.t_max), "-o", d + "/o.txt",
                   "-p", repr(plan.phi)] + (["-r"] if w.track_ratio else [])
            walls = []
            for _ in range(3):
                t0 = time.perf_counter()
                subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
                walls.append(time.perf_counter() - t0)
            res["cli_wall_s"] = walls
            res["cli_output_rows"] = len(Path(d, "o.txt").read_text().splitlines())
    out["config%d" % cfg] = res
    print(cfg, json.dumps(res), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "all_configs.json").write_text(json.dumps(out, indent=1))
