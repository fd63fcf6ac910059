#!/usr/bin/env python3
"""Which build of the reference simulates the config-2 input correctly on a B200, and how long does it take?

  gpurun -- python tools/ref_probe.py            -> gpurun_out/ref_probe.json

Binaries (oracle/Makefile): procell_ref = unmodified; procell_ref_pl* = the same with the ONE declared line
cudaDeviceSetLimit(cudaLimitDevRuntimePendingLaunchCount, N).  For every binary and input size
the leaf total and the fluorescence mass of its output are compared with the Philox oracle's expectation on the same
input (equal in law; mass is conserved exactly while phi does not bind), and the wall time is recorded next to the
process-start floor (the same command with -t 0: context creation, seed population, no division).
This script is a measurement tool of the reference arm: it uses the oracle as the checker, like bench.py --impl reference."""
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import oracle_lib  # noqa: E402
from cuda_pro_cell_b200 import synth  # noqa: E402

REFDIR = ROOT / "oracle" / "_ref"
OUT = ROOT / "gpurun_out"
TMP = OUT / "tmp_ref"
TMP.mkdir(parents=True, exist_ok=True)


def run(binary, w, t_max, limit=600):
    h, c, o = TMP / "h.txt", TMP / "c.txt", TMP / "o.txt"
    h.write_text(synth.histogram_text(w.values, w.freqs))
    c.write_text(synth.types_text(w.types[0]))
    if o.exists():
        o.unlink()
    cmd = [str(binary), "-h", str(h), "-c", str(c), "-t", repr(float(t_max)), "-p", repr(float(w.phi)), "-o", str(o)] + (["-r"] if w.track_ratio else [])
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=limit)
        rc = r.returncode
    except subprocess.TimeoutExpired:
        rc = -999
    dt = time.perf_counter() - t0
    rows = [ln.split("\t") for ln in o.read_text().splitlines()] if o.exists() else []
    return dict(rc=rc, wall_s=dt, leaves=sum(int(x[1]) for x in rows), mass=sum(float(x[0]) * int(x[1]) for x in rows))


def main():
    sizes = [int(float(x)) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["2e4", "1e5", "1e6"])]
    binaries = sorted(p for p in REFDIR.glob("procell_ref*") if p.suffix == "" and p.is_file())
    out = {"binaries": [b.name for b in binaries], "cases": []}
    for n in sizes:
        w = synth.workload(2, n / 1e6)
        oplan = oracle_lib.OraclePlan(w.values, w.freqs, w.phi)
        exp = oracle_lib.simulate(oplan, w.types, w.t_max, w.seed)
        exp_leaves, exp_div = int(exp["row_freq"].sum()), int(exp["divisions"].sum())
        exp_mass = float((exp["row_freq"][0] * oplan.row_value).sum())
        for b in binaries:
            floor = run(b, w, 0.0)
            time.sleep(1.1)
            full = run(b, w, w.t_max)
            time.sleep(1.1)
            rec = dict(binary=b.name, n_cells=n, expected_leaves=exp_leaves, expected_divisions=exp_div, expected_mass=exp_mass,
                       floor=floor, run=full, leaf_ratio=full["leaves"] / exp_leaves if exp_leaves else None,
                       mass_ratio=full["mass"] / exp_mass if exp_mass else None,
                       div_per_s_wall=exp_div / full["wall_s"], div_per_s_minus_floor=exp_div / max(full["wall_s"] - floor["wall_s"], 1e-9))
            out["cases"].append(rec)
            print(json.dumps(rec), flush=True)
            (OUT / "ref_probe.json").write_text(json.dumps(out, indent=1))
            if full["rc"] == -999:
                break
    print("done")


if __name__ == "__main__":
    main()
