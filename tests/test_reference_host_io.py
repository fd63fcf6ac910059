"""Reader, plan and writer pinned to the reference's OWN host code (no GPU, no restatement in between).

oracle/_ref/libref_host.so is the reference's translation units compiled where they lie under /root/reference plus a
C entry (oracle/ref_host_shim.cu): io::load_fluorescences (parser.cu:68-154), io::load_cell_types (parser.cu:156-185)
and io::save_fluorescences (parser.cu:187-217) are pure host code and run here.  What is compared:
  * which histogram lines are read and where reading stops, the default phi, the cell total, the seed-cell bounds of
    every bin, and the set of result rows (bit patterns of the doubles) - against procell_read_histogram +
    procell_plan_create on this side and against the oracle's plan;
  * the cell-types reader and the order the reference sorts the types into (which decides the type of a seed cell for
    a given uniform, cell.cu:81-104) - against procell_read_cell_types and the order the oracle uses;
  * the bytes io::save_fluorescences writes - against procell_write_histogram.
The library is built only where /root/reference exists (`make -C oracle refhost`); elsewhere these tests skip."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cuda_pro_cell_b200 import api, synth
import oracle_lib

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libref_host.so"

_f64p, _u64p, _i32p = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def ref():
    if not LIB.exists() and Path("/root/reference/src/io/parser.cu").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "refhost"], check=True, capture_output=True, timeout=900)
    if not LIB.exists():
        pytest.skip("reference sources not present: oracle/_ref/libref_host.so cannot be built here")
    L = C.CDLL(str(LIB))
    L.ref_load_fluorescences.argtypes = [C.c_char_p, C.c_double, _f64p, _u64p, _f64p, _u64p, _u64p, C.c_size_t,
                                         C.POINTER(C.c_size_t), _f64p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.ref_load_cell_types.argtypes = [C.c_char_p, _i32p, _f64p, _f64p, _f64p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.ref_save_fluorescences.argtypes = [C.c_char_p, C.c_int, C.c_int32, _f64p, _u64p, _i32p, C.c_size_t]
    return L


def ref_histogram(L, path, phi):
    cap, rcap = 1 << 17, 1 << 22
    v, f, b = np.zeros(cap), np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.uint64)
    rows = np.zeros(rcap)
    thr, tot, n, nr = C.c_double(), C.c_uint64(), C.c_size_t(), C.c_size_t()
    rc = L.ref_load_fluorescences(str(path).encode(), float(phi), C.byref(thr), C.byref(tot), v.ctypes.data_as(_f64p),
                                  f.ctypes.data_as(_u64p), b.ctypes.data_as(_u64p), cap, C.byref(n),
                                  rows.ctypes.data_as(_f64p), rcap, C.byref(nr))
    assert rc == 0
    return dict(phi=thr.value, total=tot.value, value=v[:n.value].copy(), freq=f[:n.value].copy(),
                bound=b[:n.value].copy(), rows=rows[:nr.value].copy())


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _histogram_texts():
    rng = np.random.default_rng(20261017)
    v, f = synth.synthetic_histogram(10_000)
    texts = {
        "synthetic_1e4": "".join("%.10g %d\n" % (x, n) for x, n in zip(v, f)),
        "readme_style": "8.144 53\n9.823 274\n11.85 1010\n14.29 2437\n17.24 3512\n",
        "unsorted_with_duplicates": "100 3\n25 7\n100 2\n12.5 1\n50 0\n3.125 9\n",
        "powers_of_two_share_rows": "1024 5\n512 4\n2 1\n1 6\n0.5 2\n",
        "zero_lines_and_tabs": "1 0\n\t2\t0\n  4  3\n8 0\n16\t5\n\n\n",
        "stops_at_garbage": "10 5\n20 6\n30 x 40 7\n50 8\n",
        "stops_at_odd_token": "10 5\n20 6\n30\n",
        "negative_frequency_wraps": "10 5\n20 -6\n30 7\n",      # operator>> reads -6 into a uint64 as 2^64 - 6
        "fraction_frequency_stops": "10 5\n20 6.5\n30 7\n",
        "exponent_and_signs": "1e3 5\n+2.5E2 6\n.5e1 7\n1.e0 8\n",
        "tiny_and_huge": "1e-300 2\n1e300 3\n4.9e-324 1\n",
        "one_line": "1009.5 12345\n",
        "only_zero_frequencies": "5 0\n6 0\n",
        "empty": "",
        "hex_and_inf_stop": "10 1\n0x10 2\ninf 3\n",
        "wide_frequency": "7 18446744073709551615\n8 1\n",
        "frequency_overflow_stops": "7 5\n8 18446744073709551616\n9 1\n",
    }
    texts["random_values"] = "".join("%.17g %d\n" % (x, n) for x, n in
                                     zip(rng.lognormal(5.0, 2.0, 300), rng.integers(0, 50, 300)))
    return texts


@pytest.mark.parametrize("name", sorted(_histogram_texts()))
@pytest.mark.parametrize("phi", [0.0, 0.5, 10.0, 1e-7])
def test_reader_and_plan_equal_load_fluorescences(ref, tmp_path, name, phi):
    text = _histogram_texts()[name]
    p = tmp_path / "h.txt"
    p.write_text(text)
    want = ref_histogram(ref, p, phi)

    v, f = api.read_histogram(p)
    keep = f > 0                                                   # parser.cu:107-108: zero-frequency lines are skipped
    assert np.array_equal(_bits(v[keep]), _bits(want["value"])) and np.array_equal(f[keep], want["freq"])
    bounds = np.zeros(int(keep.sum()), dtype=np.uint64)
    with np.errstate(over="ignore"):
        bounds[1:] = np.cumsum(f[keep], dtype=np.uint64)[:-1]
    assert np.array_equal(bounds, want["bound"])                   # parser.cu:110-126: first seed cell of every bin
    if not keep.any():
        # no cells at all: the reference's threshold stays 0 and it has no rows; this side refuses to build a plan
        assert want["total"] == 0 and len(want["rows"]) == 0
        return
    if (f >= 2 ** 32).any():
        # a frequency this build's 2^32-cell key layout cannot hold, e.g. "-6" read as 2^64 - 6: either
        # refused, or the cell total wraps exactly as the reference's uint64 sum does
        try:
            assert api.Plan(v, f, phi).n_cells == want["total"]
        except api.ProcellError:
            pass
        return
    for plan in (api.Plan(v, f, phi), oracle_lib.OraclePlan(v, f, phi)):
        assert plan.phi == want["phi"]                             # default phi: parser.cu:80-96
        assert plan.n_cells == want["total"] and plan.n_bins == len(want["value"])
        if name == "tiny_and_huge":
            # 1e300 halved down to phi: 1000-2000 levels.  This build stops at tree depth 63 and says so
            # (plan.depth_capped); its rows are the reference's rows of the first 64 levels.
            assert api.Plan(v, f, phi).depth_capped and np.isin(_bits(plan.row_value), _bits(want["rows"])).all()
            continue
        # parser.cu:128-151, std::map order.  One row of the reference's set can never be counted and is therefore
        # never written (parser.cu:198): value == phi reached by halving - a cell divides only if f/2 > phi
        # (proliferation.cu:323), so no daughter ever has f == phi.  This side leaves that row out of the key space.
        dead = want["rows"] == want["phi"] if not (want["value"] == want["phi"]).any() else np.zeros(len(want["rows"]), bool)
        assert np.array_equal(_bits(plan.row_value), _bits(want["rows"][~dead]))


def test_default_phi_scan_sees_every_line_of_the_file(ref, tmp_path):
    """parser.cu:80-96 takes the minimum over lines with frequency > 0 only, in a first pass over the whole file"""
    p = tmp_path / "h.txt"
    p.write_text("100 5\n0.001 0\n7 2\n3 0\n50 1\n")
    want = ref_histogram(ref, p, 0.0)
    v, f = api.read_histogram(p)
    assert want["phi"] == 7.0 and api.Plan(v, f, 0.0).phi == 7.0


_TYPE_TEXTS = [
    "0.53 48.33 21.6\n0.29 86.3 26.8\n0.18 -1 -1\n",
    "0.18 -1 -1\n0.29 86.3 26.8\n0.53 48.33 21.6\n",
    "0.25 1 1\n0.25 2 2\n0.25 3 3\n0.25 4 4\n",                     # ties: the order must be the stable one
    "0.2 1 1\n0.3 2 2\n0.2 3 3\n0.3 4 4\n",
    "0.40 48.33 21.6\n0.25 86.3 26.8\n0.17 24.0 6.0\n0.18 -1 -1\n",
    "1 10 2\n",
    "0.5 1e1 2.5e0 0.5 -1 -1",
    "0.125 1 1\n" * 8,
    "0.1 1 1\n0.2 2 2\n0.3 3 3\n0.4 4 4\ntrailing words\n0.5 9 9\n",
]


@pytest.mark.parametrize("text", _TYPE_TEXTS)
def test_cell_types_reader_and_sort_equal_load_cell_types(ref, tmp_path, text):
    p = tmp_path / "c.txt"
    p.write_text(text)
    cap = 64
    name, prop, mean, sd = (np.zeros(cap, dtype=np.int32), np.zeros(cap), np.zeros(cap), np.zeros(cap))
    n = C.c_size_t()
    assert ref.ref_load_cell_types(str(p).encode(), name.ctypes.data_as(_i32p), prop.ctypes.data_as(_f64p),
                                   mean.ctypes.data_as(_f64p), sd.ctypes.data_as(_f64p), cap, C.byref(n)) == 0
    n = n.value
    mine = api.read_cell_types(p)                                  # file order, type id = line index (parser.cu:162-175)
    assert mine.shape == (n, 3)
    order = name[:n]
    assert sorted(order.tolist()) == list(range(n))
    # the reference hands back the types sorted by descending proportion (parser.cu:184); entry j is file line order[j]
    assert np.array_equal(_bits(mine[order, 0]), _bits(prop[:n])) and np.array_equal(_bits(mine[order, 1]), _bits(mean[:n]))
    assert np.array_equal(_bits(mine[order, 2]), _bits(sd[:n]))
    # the selection order used by this build (capi.cu: std::stable_sort) and by the oracle (sort_types): stable, descending
    assert order.tolist() == np.argsort(-mine[:, 0], kind="stable").tolist()


def _rows(rng, n, n_types):
    special = [0.1, 1.0, 1234567890.125, 1e-5, 99999.999995, 1e15, 5e-324, 1e-310, 9.9999999995e9, 0.000099999999995,
               1e100, 2.0 ** 53, 1.0 / 3.0, 9999999999.5, 0.00001234567890123, 1.7e308]   # %.10g corner cases
    value = np.sort(np.concatenate([rng.lognormal(3.0, 3.0, n - len(special)), special]))
    freq = rng.integers(0, 5_000_000, n).astype(np.uint64)
    freq[rng.integers(0, n, n // 5)] = 0                           # rows with frequency 0 are not written
    ratio = rng.integers(0, 2_000_000, (n, max(n_types, 1))).astype(np.int32)
    return value, freq, ratio


@pytest.mark.parametrize("n_types", [0, 1, 3, 4])
def test_writer_bytes_equal_save_fluorescences(ref, tmp_path, n_types):
    rng = np.random.default_rng(7 + n_types)
    value, freq, ratio = _rows(rng, 4000, n_types)
    a, b = tmp_path / "ref.txt", tmp_path / "new.txt"
    assert ref.ref_save_fluorescences(str(a).encode(), int(n_types > 0), n_types, value.ctypes.data_as(_f64p),
                                      freq.ctypes.data_as(_u64p), ratio.ctypes.data_as(_i32p), len(value)) == 0
    api.write_histogram(str(b), value, freq.astype(np.int64), ratio.astype(np.int64) if n_types else None)
    assert a.read_bytes() == b.read_bytes() and a.stat().st_size > 40_000


def test_written_file_is_read_back_identically_by_the_reference(ref, tmp_path):
    """output of this build -> the reference's reader: a result histogram is a valid input histogram (ProCell chains
    time points that way), and %.10g is what the reference itself would have written"""
    rng = np.random.default_rng(99)
    value, freq, _ = _rows(rng, 500, 0)
    p = tmp_path / "o.txt"
    api.write_histogram(str(p), value, freq.astype(np.int64), None)
    want = ref_histogram(ref, p, 0.0)
    v, f = api.read_histogram(p)
    assert np.array_equal(_bits(v), _bits(want["value"])) and np.array_equal(f, want["freq"]) and (f > 0).all()
    assert int(f.sum()) == int(freq.sum())


def test_random_histograms_equal_load_fluorescences(ref, tmp_path):
    """300 random histogram files (values on a coarse grid so that equal values, exact halves and values equal to phi
    occur; zero frequencies; explicit and default phi) through the reference's loader and this build's reader + plan."""
    rng = np.random.default_rng(424242)
    p = tmp_path / "h.txt"
    for it in range(300):
        n = int(rng.integers(1, 40))
        style = it % 3
        if style == 0:
            vals = rng.choice([0.25, 0.5, 1.0, 2.0, 3.0, 4.0, 6.0, 8.0, 12.0, 16.0, 100.0, 1024.0], n)
        elif style == 1:
            vals = np.round(rng.lognormal(3.0, 2.0, n), 3)
        else:
            vals = rng.integers(1, 50, n).astype(np.float64) * 0.5
        freqs = rng.integers(0, 6, n) * rng.integers(0, 2, n)
        phi = [0.0, 0.5, float(vals.min()), float(vals.max()) / 4.0][int(rng.integers(0, 4))]
        p.write_text("".join("%.17g %d\n" % (v, f) for v, f in zip(vals, freqs)))
        want = ref_histogram(ref, p, phi)
        v, f = api.read_histogram(p)
        keep = f > 0
        assert np.array_equal(_bits(v[keep]), _bits(want["value"])) and np.array_equal(f[keep], want["freq"])
        if not keep.any() or phi > float(v[keep].max()) * 2:
            continue
        plan = api.Plan(v, f, phi)
        assert plan.phi == want["phi"] and plan.n_cells == want["total"] and plan.n_bins == len(want["value"])
        dead = want["rows"] == want["phi"] if not (want["value"] == want["phi"]).any() else np.zeros(len(want["rows"]), bool)
        assert np.array_equal(_bits(plan.row_value), _bits(want["rows"][~dead])), (it, vals.tolist(), freqs.tolist(), phi)
        # every (bin, level) key points at the row that holds value / 2^level
        kb, kd = plan.bin_keybase, plan.bin_kdiv & 63
        for b, val in enumerate(v[keep]):
            for k in range(int(kd[b]) + 1):
                row = plan.key_row[kb[b] + k]
                if row != 0xFFFFFFFF:
                    assert plan.row_value[row] == val / 2.0 ** k
