#!/usr/bin/env python3
"""Golden count tensors of the BASELINE configs the CPU oracle cannot follow inside a test's time limit.

  python tests/golden/make_full_size_golden.py 4        -> tests/golden/oracle_config4_full.npz

Config 4 (1e4 seed cells, t_max 720, phi 1e-7: 8.1e10 divisions) takes the oracle several minutes on all host cores;
it is run ONCE here (CPU only, oracle/liboracle.so) and the whole int64 tensor [n_keys][n_types] is committed
(compressed: most keys are empty), so that tests/test_gpu_parity.py can compare the GPU's full-size tensor bit for
bit.  The file also records the oracle's division total, the inputs' identity (config, seed, n_keys) and the wall time.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib  # noqa: E402
from cuda_pro_cell_b200 import synth  # noqa: E402


def main():
    config = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    w = synth.workload(config)
    plan = oracle_lib.OraclePlan(w.values, w.freqs, w.phi)
    t0 = time.time()
    r = oracle_lib.simulate(plan, w.types, w.t_max, w.seed)
    dt = time.time() - t0
    out = Path(__file__).resolve().parent / ("oracle_config%d_full.npz" % config)
    np.savez_compressed(out, counts=r["counts"], divisions=r["divisions"], seed=np.uint64(w.seed), n_keys=np.int64(plan.n_keys),
                        t_max=np.float64(w.t_max), phi=np.float64(plan.phi), n_cells=np.int64(plan.n_cells),
                        oracle_wall_s=np.float64(dt), oracle_threads=np.int64(oracle_lib.n_host_threads()))
    print("config %d: %d divisions, %d leaves, %.1f s on %d threads -> %s (%d bytes)"
          % (config, int(r["divisions"].sum()), int(r["counts"].sum()), dt, oracle_lib.n_host_threads(), out, out.stat().st_size))


if __name__ == "__main__":
    main()
