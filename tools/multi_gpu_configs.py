#!/usr/bin/env python3
"""BASELINE's own multi-GPU configurations through the single-process path (procell_proliferate_multi: one engine per
GPU, seed-cell units or subtrees sharded over them, ONE ncclReduce(sum, int64) onto GPU 0):

  gpurun --gpus 8 -- python tools/multi_gpu_configs.py 8      -> gpurun_out/multi_gpu_configs_8gpu.json

  config 3  1e8 seed cells, default phi, t_max 336: strong scaling over 1 / 2 / 4 / 8 GPUs
  config 4  deep trees (1e4 cells, phi 1e-7, t_max 720): all GPUs, lineage sharding (level 0) against subtree sharding at
            tree levels 4 / 6 / 8
Every reduced tensor is compared with the single-GPU tensor of the same run (bit for bit); times are the library's own
CUDA-event span over launch + reduce on GPU 0 (kernel_ms), communicator set-up excluded."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
from cuda_pro_cell_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
out = {"n_gpus_of_the_box": n}

w = synth.workload(3)
plan = api.Plan(w.values, w.freqs, w.phi)
single = api.proliferate(plan, w.types, w.t_max, w.seed)
res = {}
for g in sorted({1, 2, n // 2, n} - {0}):
    best = None
    for _ in range(2):
        t = time.time()
        r = api.proliferate_multi(plan, w.types, w.t_max, w.seed, n_gpus=g)
        dt = time.time() - t
        best = r.stats["kernel_ms"] if best is None else min(best, r.stats["kernel_ms"])
    res[g] = dict(kernel_ms=best, wall_s_incl_setup=dt, divisions=int(r.divisions.sum()),
                  equals_single_gpu=bool(np.array_equal(r.counts, single.counts) and np.array_equal(r.divisions, single.divisions)))
    print("config 3 on", g, "GPU(s):", res[g], flush=True)
base = res[1]["kernel_ms"]
for g in res:
    res[g]["speedup"] = base / res[g]["kernel_ms"]
    res[g]["efficiency"] = base / res[g]["kernel_ms"] / g
out["config3_strong_scaling"] = res

w = synth.workload(4)
plan = api.Plan(w.values, w.freqs, w.phi)
single = api.proliferate(plan, w.types, w.t_max, w.seed)
res = {"single_gpu_kernel_ms": single.stats["kernel_ms"]}
for level in (0, 4, 6, 8):
    best = None
    for _ in range(2):
        r = api.proliferate_multi(plan, w.types, w.t_max, w.seed, n_gpus=n, shard_level=level)
        best = r.stats["kernel_ms"] if best is None else min(best, r.stats["kernel_ms"])
    res["level%d" % level] = dict(kernel_ms=best, divisions=int(r.divisions.sum()), speedup=single.stats["kernel_ms"] / best,
                                  efficiency=single.stats["kernel_ms"] / best / n,
                                  equals_single_gpu=bool(np.array_equal(r.counts, single.counts) and np.array_equal(r.divisions, single.divisions)))
    print("config 4 on", n, "GPUs, shard level", level, ":", res["level%d" % level], flush=True)
out["config4_%dgpu" % n] = res
json.dump(out, open(ROOT / "gpurun_out" / ("multi_gpu_configs_%dgpu.json" % n), "w"), indent=1)
