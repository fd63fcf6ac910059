"""Known answers derivable from the reference source, asserted on the CPU oracle.
The same properties are asserted on the GPU kernels in test_gpu_parity.py and on the reference binary's own
output fixtures in test_reference_fixtures.py."""
import numpy as np
import pytest

from cuda_pro_cell_b200 import synth


def _plan(oracle, n=3000, phi=1.0):
    v, f = synth.synthetic_histogram(n)
    return v, f, oracle.OraclePlan(v, f, phi)


def test_tmax_zero_is_identity(oracle):
    # every proliferating seed has t0 + timer0 > 0 -> out_of_time at level 0 (proliferation.cu:408-409)
    v, f, p = _plan(oracle)
    r = oracle.simulate(p, [synth.TYPES_CONFIG1], 0.0, 1)
    rf = r["row_freq"][0]
    assert int(r["divisions"][0]) == 0
    assert np.array_equal(rf[rf > 0], f[f > 0].astype(np.int64))
    assert np.array_equal(p.row_value[rf > 0], v[f > 0])


def test_all_quiescent_is_identity(oracle):
    v, f, p = _plan(oracle)
    r = oracle.simulate(p, [[(1.0, -1.0, -1.0)]], 1000.0, 2)
    rf = r["row_freq"][0]
    assert np.array_equal(rf[rf > 0], f[f > 0].astype(np.int64)) and int(r["divisions"][0]) == 0


def test_mass_and_count_conservation_when_phi_never_binds(oracle):
    v, f, p = _plan(oracle, 4000, 1e-9)
    r = oracle.simulate(p, [synth.TYPES_CONFIG1], 168.0, 3)
    rf = r["row_freq"][0]
    assert int(rf.sum()) - p.n_cells == int(r["divisions"][0])            # each division adds exactly one leaf
    assert np.isclose(float((rf * p.row_value).sum()), float((v * f).sum()), rtol=1e-12)   # halving conserves mass


def test_phi_drop_loses_cells_silently(oracle):
    # alive, in time, f/2 <= phi -> neither divided nor counted
    v, f, p = _plan(oracle, 4000, 0.0)          # default phi = smallest non-empty value
    assert p.phi == float(v[f > 0].min())
    r = oracle.simulate(p, [synth.TYPES_CONFIG1], 400.0, 3)
    assert int(r["row_freq"][0].sum()) < p.n_cells + int(r["divisions"][0])


def test_ratio_columns_sum_to_total_and_follow_file_order(oracle):
    v, f, p = _plan(oracle, 20000, 1.0)
    types = [(0.2, -1.0, -1.0), (0.5, 48.33, 21.6), (0.3, 86.3, 26.8)]      # quiescent FIRST in the file
    r = oracle.simulate(p, [types], 168.0, 4)
    assert np.array_equal(r["row_ratio"][0].sum(axis=1), r["row_freq"][0])
    q = r["counts"][0][:, 0]                       # file index 0 = the quiescent type: only level-0 keys
    assert int(q.sum()) == int(q[p.bin_keybase].sum())
    assert abs(int(q.sum()) - 0.2 * 20000) < 5 * np.sqrt(20000 * 0.2 * 0.8)   # share = proportion (binomial CI)


def test_sigma_zero_closed_form(oracle):
    """sd = 0: every timer equals the mean, so a seed with initial age t0 yields 2^k leaves at level k with
    k = #{j >= 1 : t0 + j*mean <= t_max} by repeated double addition (phi never binds here)."""
    v = np.array([1024.0])
    f = np.array([500], dtype=np.uint64)
    p = oracle.OraclePlan(v, f, 1e-6)
    mean, t_max = 30.0, 100.0
    r = oracle.simulate(p, [[(1.0, mean, 0.0)]], t_max, 5)
    counts = r["counts"][0][:, 0]
    # t0 = mean * U in (0, mean): k = 3 if t0 + 3*mean <= 100 i.e. U <= 1/3, else 2
    assert set(np.nonzero(counts)[0]) <= {2, 3}
    assert counts[2] % 4 == 0 and counts[3] % 8 == 0
    n3 = counts[3] // 8
    assert counts[2] // 4 + n3 == 500
    assert abs(n3 - 500 / 3) < 5 * np.sqrt(500 * (1 / 3) * (2 / 3))


def test_sigma_zero_tree_closed_form_at_depth(oracle):
    """the same closed form as a shared checker (tests/closed_forms.py); the GPU runs it at depth 33, where single
    counts exceed 2^32 (test_gpu_parity.py)"""
    from closed_forms import check_sigma_zero_tree

    def sim(values, freqs, phi, types, t_max):
        p = oracle.OraclePlan(values, freqs, phi)
        r = oracle.simulate(p, types, t_max, 5)
        return r["counts"][0], r["divisions"][0], int(p.bin_keybase[0])

    for n_seeds, k_hi in ((6, 12), (3, 15), (40, 5)):
        check_sigma_zero_tree(sim, n_seeds, k_hi)


def test_key_space_and_row_merge(oracle):
    # 8 and 4 share the rows 4, 2, 1: equal value/2^k from different (bin, k) merge into one output row
    p = oracle.OraclePlan(np.array([8.0, 4.0, 3.0]), np.array([1, 1, 1], dtype=np.uint64), 1.0)
    assert p.bin_kdiv.tolist() == [2, 1, 1]          # f/2 > phi strictly: 8 -> 4 -> 2 (2/2 = 1 is not > 1)
    assert p.row_value.tolist() == [1.5, 2.0, 3.0, 4.0, 8.0]
    assert p.key_row.tolist() == [4, 3, 1, 3, 1, 2, 0]
    # a bin below phi: no countable level, no division
    p2 = oracle.OraclePlan(np.array([0.5, 8.0]), np.array([3, 1], dtype=np.uint64), 1.0)
    assert p2.bin_kdiv.tolist() == [0, 2] and p2.key_row[0] == 0xFFFFFFFF
    # zero-frequency lines are skipped, duplicates allowed, order preserved
    p3 = oracle.OraclePlan(np.array([5.0, 9.0, 5.0]), np.array([2, 0, 1], dtype=np.uint64), 1.0)
    assert p3.n_bins == 2 and p3.n_cells == 3


def test_shards_partition_the_run(oracle):
    v, f, p = _plan(oracle, 5000, 0.5)
    whole = oracle.simulate(p, [synth.TYPES_CONFIG2], 200.0, 6)
    for world, unit in ((2, 256), (3, 7), (8, 32)):
        parts = [oracle.simulate(p, [synth.TYPES_CONFIG2], 200.0, 6, shard=(r, world, unit)) for r in range(world)]
        assert np.array_equal(sum(x["counts"] for x in parts), whole["counts"])
        assert sum(int(x["divisions"][0]) for x in parts) == int(whole["divisions"][0])


def test_subtree_shards_partition_the_run_and_balance_deep_trees(oracle):
    """Subtree sharding (oracle_simulate, shard_level >= 1): every rank walks the first levels of every
    lineage, a subtree at the shard level belongs to one rank.  The ranks' tensors sum to the unsharded run for any
    level and world size - also when the level is deeper than the trees - and on the config-4 shape (1 % of the seed
    cells own almost all divisions) the busiest rank carries a few per cent more than the mean instead of ~40 %."""
    v, f, p = _plan(oracle, 3000, 0.5)
    whole = oracle.simulate(p, [synth.TYPES_CONFIG2], 200.0, 6)
    for world, level in ((2, 1), (3, 2), (8, 5), (5, 40)):
        parts = [oracle.simulate(p, [synth.TYPES_CONFIG2], 200.0, 6, shard=(r, world, 32), shard_level=level) for r in range(world)]
        assert np.array_equal(sum(x["counts"] for x in parts), whole["counts"])
        assert sum(int(x["divisions"][0]) for x in parts) == int(whole["divisions"][0])
        assert all(int(x["divisions"][0]) > 0 for x in parts)
    v, f = synth.synthetic_histogram(10000)
    p = oracle.OraclePlan(v, f, 1e-7)
    whole = oracle.simulate(p, [synth.TYPES_CONFIG4], 300.0, 0x5EED0004)

    def busiest_over_mean(level):
        parts = [oracle.simulate(p, [synth.TYPES_CONFIG4], 300.0, 0x5EED0004, shard=(r, 8, 32), shard_level=level) for r in range(8)]
        assert np.array_equal(sum(x["counts"] for x in parts), whole["counts"])
        d = [int(x["divisions"][0]) for x in parts]
        assert sum(d) == int(whole["divisions"][0])
        return max(d) * 8 / sum(d)

    by_lineage, by_subtree = busiest_over_mean(0), busiest_over_mean(6)
    assert by_lineage > 1.25 and by_subtree < 1.05, (by_lineage, by_subtree)


def test_seed_and_set_change_the_stream_but_not_the_law(oracle):
    v, f, p = _plan(oracle, 20000, 1.0)
    a = oracle.simulate(p, [synth.TYPES_CONFIG1], 168.0, 100)
    b = oracle.simulate(p, [synth.TYPES_CONFIG1], 168.0, 101)
    assert not np.array_equal(a["counts"], b["counts"])
    assert abs(int(a["divisions"][0]) - int(b["divisions"][0])) < 0.05 * int(a["divisions"][0])
    two = oracle.simulate(p, [synth.TYPES_CONFIG1, synth.TYPES_CONFIG1], 168.0, 100)   # same parameters, two sets
    assert np.array_equal(two["counts"][0], a["counts"][0])
    assert not np.array_equal(two["counts"][1], a["counts"][0])


def test_refcompat_couples_type_and_first_timer(oracle):
    """in the reference the type uniform is also the Box-Muller radius uniform of the first timer, which
    shifts the generation shares of each type; the refcompat mode reproduces that coupling, the ideal mode does not."""
    v, f = synth.synthetic_histogram(100000)
    p = oracle.OraclePlan(v, f, 1e-9)
    types = [synth.TYPES_CONFIG1]

    def shares(ref):
        r = oracle.simulate(p, types, 168.0, 9, refcompat=ref)
        c = r["counts"][0][:, 1].astype(float)          # type 1 = (0.29, 86.3, 26.8)
        k = np.concatenate([np.arange(n + 1) for n in p.bin_kdiv])
        by_k = np.bincount(k, weights=c, minlength=8)
        return by_k / by_k.sum()
    ideal, ref = shares(False), shares(True)
    # survey: type-1 leaf share at k=1 is ~0.54 with independent draws and ~0.65 with the coupling
    assert 0.50 < ideal[1] < 0.58
    assert 0.61 < ref[1] < 0.69


def test_forced_timer_after_255_rejections(oracle):
    # mean tiny and sd = 0 -> never positive: every redraw is rejected, the 256th falls back to the mean (= 0 here)
    p = oracle.OraclePlan(np.array([4.0]), np.array([2], dtype=np.uint64), 1.0)
    r = oracle.simulate(p, [[(1.0, 0.0, 0.0)]], 10.0, 1)
    # timer 0 forever: cells divide instantly until phi stops them (kdiv = 1), daughters are dropped
    assert int(r["divisions"][0]) == 2 and int(r["row_freq"][0].sum()) == 0


def test_proportion_check(oracle):
    L = oracle.lib()
    import ctypes as C
    ok = np.array([[0.53, 1, 1], [0.29, 1, 1], [0.18, 1, 1]], dtype=np.float64)
    bad = np.array([[0.5, 1, 1], [0.3, 1, 1]], dtype=np.float64)
    f64p = C.POINTER(C.c_double)
    assert L.oracle_check_proportions(ok.ctypes.data_as(f64p), 3) == 0
    assert L.oracle_check_proportions(bad.ctypes.data_as(f64p), 2) == 1
