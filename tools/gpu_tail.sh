#!/bin/bash
# donation-policy A/B: run stats (kernel time, seed phase, idle warp time) of config 2 / config 4 for library variants
cd "$(dirname "$0")/.."
for lib in "$@"; do
  for rep in 1 2; do
    PROCELL_LIB=$lib python - <<PY
import sys
sys.path.insert(0, '.')
from cuda_pro_cell_b200 import api, synth
out = []
for cfg, scale, tmax in ((2, 1.0, 0), (4, 0.1, 600.0), (3, 0.1, 0)):
    w = synth.workload(cfg, scale)
    if tmax: w.t_max = tmax
    plan = api.Plan(w.values, w.freqs, w.phi)
    eng = api.Engine(0); eng.load(plan, w.types, w.t_max, w.seed)
    for _ in range(3):
        eng.run(); r = eng.finish(fetch=False)
    s = r.stats
    nw = s['grid'] * s['block'] // 32
    out.append("cfg%d %.3f ms seed %.0f total %.0f us idle %.1f%% (tail idle %.0f%%) waits %d don %d" % (
        cfg, s['kernel_ms'], s['seed_phase_us'], s['total_us'], 100 * s['idle_warp_us'] / (nw * s['total_us']),
        100 * s['idle_warp_us'] / (nw * max(1e-9, s['total_us'] - s['seed_phase_us'])), s['idle_waits'], s['donations']))
print("$lib rep$rep | " + " | ".join(out))
PY
  done
done
