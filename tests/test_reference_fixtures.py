"""Pinning against the reference itself.  tests/golden/ref_*.json are outputs of the UNMODIFIED reference binary run on
a B200 (tests/golden/make_ref_fixtures.py).  The reference seeds cuRAND from time(NULL), so each fixture records the
wall-clock window of its run; oracle/xorwow_ref.c - the restatement of the reference AS WRITTEN - must reproduce every
fixture bit for bit for one (T0, T1) in that window.  The Philox oracle shares those semantics and is then compared with
the reference distributionally (chi-square and KS on the output histogram, thresholds calibrated on reference-vs-
reference splits and stated below)."""
import json
from pathlib import Path

import numpy as np
import pytest

from cuda_pro_cell_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"


def _load(name):
    return json.loads((GOLDEN / ("ref_%s.json" % name)).read_text())


def _rows_of(sim, track_ratio=True):
    nz = sim["row_freq"] > 0
    out = []
    for v, f, r in zip(sim["row_value"][nz], sim["row_freq"][nz], sim["row_ratio"][nz]):
        out.append(["%.10g" % v, int(f)] + ([int(x) for x in r] if track_ratio else []))
    return out


@pytest.mark.parametrize("name", ["cfg1", "cfg1_phi_tiny", "cfg1_tmax0", "cfg1_quiescent", "cfg2_2k"])
def test_xorwow_restatement_reproduces_the_reference_binary(oracle, name):
    fx = _load(name)
    values, freqs = synth.synthetic_histogram(fx["n_cells"])
    for run in fx["runs"]:
        assert run["rc"] == 0
        lo, hi = run["time_window"]
        hit = None
        for T0 in range(lo, hi + 1):
            for T1 in (T0, T0 + 1):        # time(NULL) is read again at the first run_iteration
                sim = oracle.xorwow_simulate(values, freqs, fx["types"], fx["t_max"], fx["phi"], T0, T1)
                if sim is not None and _rows_of(sim) == run["rows"]:
                    hit = (T0, T1)
                    break
            if hit:
                break
        assert hit is not None, "no wall-clock seed in %s reproduces the reference output of %s" % (run["time_window"], name)


def test_reference_fixtures_obey_the_source_derived_invariants():
    # t_max = 0 and all-quiescent: output == input (rows with frequency > 0, same order)
    for name in ("cfg1_tmax0", "cfg1_quiescent"):
        fx = _load(name)
        values, freqs = synth.synthetic_histogram(fx["n_cells"])
        rows = fx["runs"][0]["rows"]
        assert [r[0] for r in rows] == ["%.10g" % v for v in values[freqs > 0]]
        assert [r[1] for r in rows] == [int(f) for f in freqs[freqs > 0]]
    # phi tiny: fluorescence mass conserved, leaves - seeds == divisions, ratio columns sum to the total, ascending
    fx = _load("cfg1_phi_tiny")
    for run in fx["runs"]:
        vals = np.array([float(r[0]) for r in run["rows"]])
        tot = np.array([r[1] for r in run["rows"]])
        assert np.isclose(float((vals * tot).sum()), fx["input_mass"], rtol=1e-8)   # values are printed with 10 digits
        assert all(sum(r[2:]) == r[1] for r in run["rows"])
        assert (np.diff(vals) > 0).all()


def test_reference_loses_subtrees_beyond_small_inputs(oracle):
    """Documented defect: on sm_100 the reference's dynamic-parallelism recursion silently drops
    subtrees once a level is wide.  At 2000 cells its leaf total agrees with the expectation; from 20000 cells on it
    does not (the fixtures record 25 %, 18 % and 7 % of the expected leaves at 2e4, 1e5 and 1e6 cells)."""
    w = synth.workload(2, 0.002)
    plan = oracle.OraclePlan(w.values, w.freqs, w.phi)
    per_cell = oracle.simulate(plan, w.types, w.t_max, 1)["row_freq"].sum() / 2000.0
    ok = _load("cfg2_2k")
    for run in ok["runs"]:
        assert abs(run["total"] / 2000.0 - per_cell) < 0.1 * per_cell
    for name, n in (("cfg2_20k", 2e4), ("cfg2_100k", 1e5), ("cfg2_1m", 1e6)):
        for run in _load(name)["runs"]:
            assert run["total"] / n < 0.5 * per_cell


# ---------------------------------------------------------------------------------------------------------------
def _type_k_table(values, freqs, phi, row_value, row_ratio):
    """pooled leaf counts by (type, number of halvings k): k = log2(v_bin / row value) is exact here because the
    synthetic channel values never collide across bins"""
    bins = values[freqs > 0]
    lookup = {}
    for b in bins:
        c, k = float(b), 0
        while c >= phi:
            lookup.setdefault(c, k)
            c /= 2
            k += 1
    ks = np.array([lookup[float(v)] for v in row_value])
    kmax = int(ks.max()) + 1
    tab = np.zeros((row_ratio.shape[1], kmax))
    for t in range(row_ratio.shape[1]):
        tab[t] = np.bincount(ks, weights=row_ratio[:, t], minlength=kmax)
    return tab


def _chi2(a, b):
    """two-sample chi-square statistic per degree of freedom over cells with enough mass"""
    a, b = a.ravel().astype(float), b.ravel().astype(float)
    keep = (a + b) >= 50
    a, b = a[keep], b[keep]
    ka, kb = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    return float((((ka * a - kb * b) ** 2) / (a + b)).sum() / max(1, keep.sum() - 1))


def _ks(a, b):
    fa, fb = np.cumsum(a) / a.sum(), np.cumsum(b) / b.sum()
    return float(np.abs(fa - fb).max())


def test_philox_oracle_agrees_with_the_reference_in_law(oracle):
    """Config 1 (phi tiny so that nothing is dropped), R = 24 pooled runs per side.
    Statistic 1: chi-square per degree of freedom between pooled (type, k) leaf tables.
    Statistic 2: two-sample KS distance between the pooled output histograms (CDF over ascending fluorescence).
    Leaves are clustered by lineage, so the thresholds are calibrated empirically: the reference-vs-reference null is
    sampled from 12 random splits of 48 emulated reference runs, and the bar is 2x the largest null value
    (measured: null chi2/dof max 4.3, median 1.8, null KS max 0.0040; refcompat scores 2.0 and 0.0033, ideal
    seeding scores 1434 and 0.053).
    refcompat seeding must pass both; ideal seeding must FAIL the chi-square (power check)."""
    R = 24
    w = synth.workload(1)
    phi = 1e-6
    values, freqs = w.values, w.freqs
    types = w.types[0]
    plan = oracle.OraclePlan(values, freqs, phi)

    refs = []
    for i in range(2 * R):
        s = oracle.xorwow_simulate(values, freqs, types, w.t_max, phi, 1_700_000_000 + 7 * i)
        refs.append((s["row_freq"], s["row_ratio"]))
    row_value = s["row_value"]
    assert np.array_equal(row_value, plan.row_value)

    def pooled(items):
        return sum(x[0] for x in items), sum(x[1] for x in items)

    def philox(refcompat):
        out = []
        for i in range(R):
            r = oracle.simulate(plan, [types], w.t_max, 1000 + i, refcompat=refcompat)
            out.append((r["row_freq"][0], r["row_ratio"][0]))
        return pooled(out)

    rng = np.random.default_rng(0)
    null_chi, null_ks = [], []
    for _ in range(12):
        perm = rng.permutation(2 * R)
        fa, ra = pooled([refs[i] for i in perm[:R]])
        fb, rb = pooled([refs[i] for i in perm[R:]])
        null_chi.append(_chi2(_type_k_table(values, freqs, phi, row_value, ra), _type_k_table(values, freqs, phi, row_value, rb)))
        null_ks.append(_ks(fa, fb))
    chi_bar, ks_bar = 2.0 * max(null_chi), 2.0 * max(null_ks)

    f_ref, r_ref = pooled(refs[:R])
    t_ref = _type_k_table(values, freqs, phi, row_value, r_ref)
    f_new, r_new = philox(True)
    chi_new = _chi2(_type_k_table(values, freqs, phi, row_value, r_new), t_ref)
    assert chi_new < chi_bar, (chi_new, chi_bar)
    assert _ks(f_new, f_ref) < ks_bar
    f_id, r_id = philox(False)
    chi_ideal = _chi2(_type_k_table(values, freqs, phi, row_value, r_id), t_ref)
    assert chi_ideal > 5 * chi_bar, (chi_ideal, chi_bar)     # independent draws are NOT the reference's law (Q1)
