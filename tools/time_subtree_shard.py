#!/usr/bin/env python3
"""Kernel time of ONE rank's share of BASELINE config 4 under subtree sharding (rank 3 of 8, level 6) on one GPU:
  [PROCELL_LIB=libprocell_b200_x.so] python tools/time_subtree_shard.py      (A/B of library builds of the MODE 1 instance)"""
import sys, os; sys.path.insert(0, '.')
from cuda_pro_cell_b200 import api, synth
w = synth.workload(4)
plan = api.Plan(w.values, w.freqs, w.phi)
best = 1e9
for i in range(3):
    r = api.proliferate(plan, w.types, w.t_max, w.seed, shard=(3, 8, 32), shard_level=6)
    best = min(best, r.stats["kernel_ms"])
print(os.environ.get("PROCELL_LIB", "main"), "config 4, rank 3 of 8, subtree level 6: %.3f ms, %d divisions" % (best, int(r.divisions[0])))
