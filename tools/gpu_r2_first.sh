#!/bin/bash
# First GPU visit after a session without GPU access (round 1 ended with the GPU budget spent).
#   gpurun --timeout 900 -- 'bash tools/gpu_r2_first.sh'            (1 GPU)
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_r2_first.sh 8'  (adds the multi-GPU legs)
# 1. the regular parity suite (the 16 kernel instances of round 1 are byte-identical, tools/sass_same.py)
# 2. the code paths that have never run on a GPU (PROCELL_TEST_NEW=1): subtree sharding
# 3. bench line; 4. with 8 GPUs: config 4 at full size, lineage sharding vs subtree sharding at level 6
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1}
export PROCELL_WATCHDOG_S=60
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2a.log
PROCELL_TEST_NEW=1 timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 150 -k "subtree or set_relative or beyond_2_to_32" > gpurun_out/pytest_gpu_r2a_new.log 2>&1; echo "pytest (new paths) rc=$?"; tail -5 gpurun_out/pytest_gpu_r2a_new.log
# config 5 (1024 sets x 1e6 cells) and a tenth of it: hashed cache against the set-relative direct table
cat > /tmp/sweep_ab.py <<'PY'
import os, sys, json; sys.path.insert(0, '.')
import numpy as np
from cuda_pro_cell_b200 import api, synth
out = {}
for scale in (0.1, 1.0):
    w = synth.workload(5, scale)
    plan = api.Plan(w.values, w.freqs, w.phi)
    for mode in ("0", "1"):
        os.environ["PROCELL_SWEEP_DIRECT"] = mode
        eng = api.Engine(0); eng.load(plan, w.types, w.t_max, w.seed)
        best = None
        for _ in range(3):
            eng.run(); r = eng.finish(fetch=True)
            best = r.stats if best is None or r.stats["kernel_ms"] < best["kernel_ms"] else best
        flat = r.counts.reshape(-1).astype(np.uint64)
        chk = int((flat * (np.arange(flat.size, dtype=np.uint64) % np.uint64(1000003) + np.uint64(1))).sum() % np.uint64(1 << 61))
        out["scale%g_direct%s" % (scale, mode)] = dict(kernel_ms=best["kernel_ms"], divisions=int(r.divisions.sum()), smem=best["smem_bytes"], checksum=chk, idle_warp_us=best["idle_warp_us"])
        print(scale, mode, out["scale%g_direct%s" % (scale, mode)], flush=True)
        eng.close()
json.dump(out, open("gpurun_out/config5_setdirect_ab.json", "w"), indent=1)
PY
timeout 300 python /tmp/sweep_ab.py 2>&1 | tail -5
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r2a.json
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-engines 2 > gpurun_out/bench_r2a_e2e2.json 2> gpurun_out/bench_r2a_e2e2.err; echo "bench (2 engines in flight) rc=$?"
python - <<'PY'
import json
for f in ("bench_r2a", "bench_r2a_e2e2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g  e2e %.4g  frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
if [ "$N" -ge 2 ]; then
cat > /tmp/sub.py <<'PY'
import sys, time, json; sys.path.insert(0, '.')
import numpy as np
from cuda_pro_cell_b200 import api, synth
n = int(sys.argv[1])
w = synth.workload(4, 1.0)
plan = api.Plan(w.values, w.freqs, w.phi)
out = {}
for level in (0, 4, 6, 8):
    api.proliferate_multi(plan, w.types, 400.0, w.seed, n_gpus=n, shard_level=level)          # warm-up (contexts, NCCL)
    t = time.time(); r = api.proliferate_multi(plan, w.types, w.t_max, w.seed, n_gpus=n, shard_level=level); dt = time.time() - t
    out[level] = dict(wall_s=dt, kernel_ms=r.stats["kernel_ms"], divisions=int(r.divisions.sum()), checksum=int((r.counts * np.arange(1, r.counts.size + 1).reshape(r.counts.shape) % 1000003).sum()))
    print(level, out[level], flush=True)
json.dump(out, open("gpurun_out/config4_subtree_%dgpu.json" % n, "w"), indent=1)
# BASELINE configs[2]: 1e8 seed cells sharded over the GPUs of the box (strong scaling of one run, one ncclReduce)
w = synth.workload(3, 1.0)
plan = api.Plan(w.values, w.freqs, w.phi)
res = {}
for g in sorted({1, 2, n // 2, n} - {0}):
    api.proliferate_multi(plan, w.types, w.t_max, w.seed, n_gpus=g)
    t = time.time(); r = api.proliferate_multi(plan, w.types, w.t_max, w.seed, n_gpus=g); dt = time.time() - t
    res[g] = dict(wall_s=dt, kernel_ms=r.stats["kernel_ms"], divisions=int(r.divisions.sum()))
    print("config 3 on", g, "GPU(s):", res[g], flush=True)
json.dump(res, open("gpurun_out/config3_strong_scaling_%dgpu.json" % n, "w"), indent=1)
PY
timeout 500 python /tmp/sub.py $N 2>&1 | tail -12
fi
