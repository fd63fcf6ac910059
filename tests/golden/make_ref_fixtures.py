#!/usr/bin/env python3
"""Generate tests/golden/ref_*.json: outputs of the UNMODIFIED reference binary (oracle/_ref/procell_ref, built by
`make -C oracle ref` from /root/reference) run on a B200, together with the wall-clock window of each run (the
reference seeds its RNG from time(NULL), proliferation.cu:242 / cells_population.cu:34).  Run on the GPU box:
    gpurun -- python tests/golden/make_ref_fixtures.py   (writes gpurun_out/golden/*.json; copy into tests/golden/)
"""
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from cuda_pro_cell_b200 import synth  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "procell_ref"
OUT = ROOT / "gpurun_out" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
TMP = ROOT / "gpurun_out" / "tmp_ref"
TMP.mkdir(parents=True, exist_ok=True)

CASES = [
    # name, n_cells, types, t_max, phi (None = min non-empty bin), track_ratio, repeats, keep_rows
    ("cfg1", 10000, synth.TYPES_CONFIG1, 168.0, None, True, 4, True),
    ("cfg1_phi_tiny", 10000, synth.TYPES_CONFIG1, 168.0, 1e-6, True, 3, True),
    ("cfg1_tmax0", 10000, synth.TYPES_CONFIG1, 0.0, 1.0, True, 1, True),
    ("cfg1_quiescent", 10000, [(1.0, -1.0, -1.0)], 500.0, 1.0, True, 1, True),
    ("cfg2_2k", 2000, synth.TYPES_CONFIG2, 240.0, 0.5, True, 2, True),
    ("cfg2_20k", 20000, synth.TYPES_CONFIG2, 240.0, 0.5, True, 2, True),
    ("cfg2_100k", 100000, synth.TYPES_CONFIG2, 240.0, 0.5, True, 1, False),
    ("cfg2_1m", 1000000, synth.TYPES_CONFIG2, 240.0, 0.5, True, 1, False),
    ("cfg1_100k_t100", 100000, synth.TYPES_CONFIG1, 100.0, 1e-6, True, 1, False),
]


def main():
    for name, n, types, t_max, phi, ratio, reps, keep in CASES:
        values, freqs = synth.synthetic_histogram(n)
        if phi is None:
            phi = float(values[freqs > 0].min())
        h, c, o = TMP / "h.txt", TMP / "c.txt", TMP / "o.txt"
        h.write_text(synth.histogram_text(values, freqs))
        c.write_text(synth.types_text(types))
        runs = []
        for rep in range(reps):
            if o.exists():
                o.unlink()
            cmd = [str(REF), "-h", str(h), "-c", str(c), "-t", repr(t_max), "-p", repr(phi), "-o", str(o)] + (["-r"] if ratio else [])
            t0 = time.time()
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
                rc, err = r.returncode, (r.stdout + r.stderr)[-400:]
            except subprocess.TimeoutExpired:
                rc, err = -999, "timeout"
            t1 = time.time()
            rows = [ln.split("\t") for ln in o.read_text().splitlines()] if o.exists() else []
            total = sum(int(x[1]) for x in rows)
            mass = sum(float(x[0]) * int(x[1]) for x in rows)
            run = dict(rc=rc, wall_s=t1 - t0, time_window=[int(t0), int(t1) + 1], n_rows=len(rows), total=total, mass=mass, log=err)
            if keep:
                run["rows"] = [[x[0]] + [int(y) for y in x[1:]] for x in rows]
            runs.append(run)
            print(name, rep, {k: v for k, v in run.items() if k != "rows"}, flush=True)
            time.sleep(1.1)
        fixture = dict(name=name, n_cells=n, types=[list(t) for t in types], t_max=t_max, phi=phi, track_ratio=ratio,
                       input_mass=float((values * freqs).sum()), histogram="cuda_pro_cell_b200.synth.synthetic_histogram(n_cells)",
                       command="procell_ref -h H -c C -t t_max -p phi -o O -r", runs=runs)
        (OUT / ("ref_%s.json" % name)).write_text(json.dumps(fixture))
    print("done")


if __name__ == "__main__":
    main()
