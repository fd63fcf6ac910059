import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from cuda_pro_cell_b200 import api, synth
import oracle_lib
n_sets, cells = int(sys.argv[1]), int(sys.argv[2])
values, freqs = synth.synthetic_histogram(cells)
types = synth.sweep_types(1024)[:: 1024 // n_sets]
plan = api.Plan(values, freqs, 0.5)
got = api.proliferate(plan, types, 168.0, 0x5EED0005)
op = oracle_lib.OraclePlan(values, freqs, 0.5)
want = oracle_lib.simulate(op, types, 168.0, 0x5EED0005)
print(n_sets, cells, os.environ.get("PROCELL_COOP_WARPS"), "match", np.array_equal(got.counts, want["counts"]), got.stats)
