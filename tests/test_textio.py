"""Streaming text I/O (cuda_pro_cell_b200/csrc/textio.cpp) against an iostream restatement of the
reference's reading loops and writer (tests/textio_ref.cpp <- src/io/parser.cu:103-106, :167-175, :187-217).

The bar is byte/bit equality: the same records (bit-identical doubles, identical counts, the same stopping point on
malformed input) and the same output bytes.  No GPU involved."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cuda_pro_cell_b200 import _lib, api

ROOT = Path(__file__).resolve().parent.parent
_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)


@pytest.fixture(scope="module")
def ref():
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    so = out / "libtextio_ref.so"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(ROOT / "tests" / "textio_ref.cpp")],
                   check=True)
    R = C.CDLL(str(so))
    R.ref_read_histogram.restype = C.c_size_t
    R.ref_read_histogram.argtypes = [C.c_char_p, _f64p, _u64p, C.c_size_t]
    R.ref_read_cell_types.restype = C.c_size_t
    R.ref_read_cell_types.argtypes = [C.c_char_p, _f64p, C.c_size_t]
    R.ref_parse_histogram.restype = C.c_size_t
    R.ref_parse_histogram.argtypes = [C.c_char_p, C.c_size_t, _f64p, _u64p, C.c_size_t]
    R.ref_parse_cell_types.restype = C.c_size_t
    R.ref_parse_cell_types.argtypes = [C.c_char_p, C.c_size_t, _f64p, C.c_size_t]
    R.ref_write_histogram.restype = C.c_int
    R.ref_write_histogram.argtypes = [C.c_char_p, C.c_int, C.c_int32, C.c_size_t, _f64p, _u64p, _i64p]
    return R


def ref_parse_histogram(R, text: bytes, cap=1 << 17):
    v = np.zeros(cap, dtype=np.float64)
    f = np.zeros(cap, dtype=np.uint64)
    n = R.ref_parse_histogram(text, len(text), v.ctypes.data_as(_f64p), f.ctypes.data_as(_u64p), cap)
    assert n <= cap
    return v[:n], f[:n]


def ref_parse_types(R, text: bytes, cap=4096):
    t = np.zeros(3 * cap, dtype=np.float64)
    n = R.ref_parse_cell_types(text, len(text), t.ctypes.data_as(_f64p), cap)
    assert n <= cap
    return t[: 3 * n].reshape(-1, 3)


def parse_types_unchecked(text: bytes):
    """the product's type reader; a proportion-sum error still carries the parsed rows, so read them through the
    histogram-independent path: ProcellError on the sum is expected for random input"""
    lib = _lib.load()
    t, n = C.POINTER(_lib.CellType)(), C.c_size_t()
    rc = lib.procell_parse_cell_types(text, len(text), C.byref(t), C.byref(n))
    try:
        assert rc in (0, _lib.ERR_PROPORTION)
        return np.array([(t[i].proportion, t[i].mean, t[i].stddev) for i in range(n.value)], dtype=np.float64).reshape(-1, 3)
    finally:
        lib.procell_free(t)


def same_bits(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


CORNER_CASES = [
    b"",
    b"\n\n  \t\n",
    b"1.0 0\n8.144 53\n  9823.85\t274\n1e3 7 garbage 5\n12 3\n",
    b"1 2",                       # no trailing newline
    b"1 2 3",                     # half a record at the end
    b"+5 7\n-3.25 8\n+.5 9\n-.5e1 10\n5. 11\n",
    b"1e5 1\n1E+05 2\n1e-05 3\n1.e2 4\n.5E1 5\n",
    b"007 08\n0.000 000\n",
    b"1e 5\n2 3\n",               # bare exponent: the stream eats the 'e' and fails
    b"1e+ 5\n2 3\n",
    b"1.5e3.7 4\n",
    b"1..5 3\n",
    b"12abc 3\n",
    b"0x10 3\n",
    b"inf 3\n4 5\n",
    b"nan 3\n",
    b"-inf 3\n",
    b"1 +3\n2 4\n",               # a sign on the frequency: operator>> accepts it
    b"1 -3\n2 4\n",               # ... and a negative one wraps modulo 2^64
    b"1 3.5 7 1\n",               # frequency 3, then value .5
    b"1 18446744073709551615\n2 18446744073709551616\n3 4\n",     # max uint64, then overflow stops the loop
    b"1 99999999999999999999999\n3 4\n",
    b"1e999 3\n4 5\n",            # overflow: failbit
    b"1e-999 3\n4 5\n",           # underflow: no failbit
    b"4.9e-324 3\n2.2250738585072014e-308 1\n1.7976931348623157e308 2\n1.7976931348623159e308 9\n5 5\n",
    b"123456789012345678901234567890123456789012345678901234567890.5 3\n",
    b"0." + b"0" * 400 + b"1 3\n2 2\n",
    b"1.0000000000000002 1\n0.1 2\n0.30000000000000004 3\n9007199254740993 4\n",
    b"1\t2\r\n3\v4\f5 6\r7 8\n",
    b"1 2\x003 4\n",              # a NUL byte is not whitespace
    b"- 3\n", b". 3\n", b"-. 3\n", b"+ 3\n", b"+-5 3\n", b"e5 3\n", b"--5 3\n",
    b"1,5 3\n",
    b"5 ",
    b"5",
    b"1 2\n# comment\n3 4\n",
]


def test_histogram_reader_corner_cases_equal_the_stream(ref):
    for text in CORNER_CASES:
        v, f = api.parse_histogram(text)
        rv, rf = ref_parse_histogram(ref, text)
        assert same_bits(v, rv) and same_bits(f, rf), text


def test_cell_types_reader_corner_cases_equal_the_stream(ref):
    cases = CORNER_CASES + [
        b"0.53 48.33 21.6\n0.29 86.3 26.8\n0.18 -1 -1\n",
        b"0.5 10 2\n0.5 -1\n",                 # short last line
        b"0.5 10 2 0.5 -1 -1",
        b"1 1e1 +2\n",
    ]
    for text in cases:
        assert same_bits(parse_types_unchecked(text), ref_parse_types(ref, text)), text


def _random_number_token(rng) -> str:
    kind = rng.integers(0, 10)
    x = float(rng.standard_normal() * 10.0 ** rng.integers(-12, 13))
    if kind == 0:
        return "%.10g" % abs(x)
    if kind == 1:
        return repr(x)
    if kind == 2:
        return "%e" % x
    if kind == 3:
        return "%d" % rng.integers(0, 10**9)
    if kind == 4:
        return "+%.17g" % abs(x)
    if kind == 5:
        return ("%.6f" % abs(x)).lstrip("0") or "0"      # ".5" forms
    if kind == 6:
        return "%d." % rng.integers(0, 1000)
    if kind == 7:
        return "%.3E" % x
    if kind == 8:
        return "000%.4f" % abs(x)
    return "%.25g" % x


def test_well_formed_files_equal_the_stream(ref, tmp_path):
    rng = np.random.default_rng(11)
    seps = [" ", "\t", "  ", " \t ", "\n"]
    eols = ["\n", "\r\n", "\n\n", " \n"]
    for case in range(40):
        n = int(rng.integers(0, 400))
        parts = []
        for _ in range(n):
            parts.append(_random_number_token(rng) + seps[rng.integers(0, len(seps))] +
                         "%d" % rng.integers(0, 2**63) + eols[rng.integers(0, len(eols))])
        text = "".join(parts).encode()
        p = tmp_path / ("h%d.txt" % case)
        p.write_bytes(text)
        v, f = api.read_histogram(p)
        rv = np.zeros(n + 1)
        rf = np.zeros(n + 1, dtype=np.uint64)
        rn = ref.ref_read_histogram(str(p).encode(), rv.ctypes.data_as(_f64p), rf.ctypes.data_as(_u64p), n + 1)
        assert rn == n == len(v)
        assert same_bits(v, rv[:n]) and same_bits(f, rf[:n])
        v2, f2 = api.parse_histogram(text)
        assert same_bits(v, v2) and same_bits(f, f2)


def test_fuzzed_text_equals_the_stream(ref):
    """random strings over the alphabet numbers are made of: wherever the stream stops, the scanner stops"""
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"0123456789012345678901234567890123456789+-..eE  \t\n\nxina", dtype=np.uint8)
    for _ in range(6000):
        text = alphabet[rng.integers(0, len(alphabet), int(rng.integers(0, 48)))].tobytes()
        v, f = api.parse_histogram(text)
        rv, rf = ref_parse_histogram(ref, text, cap=64)
        assert same_bits(v, rv) and same_bits(f, rf), text
        assert same_bits(parse_types_unchecked(text), ref_parse_types(ref, text, cap=64)), text


def test_reader_at_the_largest_supported_size(ref, tmp_path):
    """65 535 lines (the key layout's limit) with zero-frequency lines in between"""
    rng = np.random.default_rng(3)
    n = 65535
    vals = np.sort(rng.uniform(1.0, 1e5, n))
    freqs = rng.integers(0, 5000, n)
    freqs[rng.integers(0, n, 5000)] = 0
    text = "".join("%.10g %d\n" % (a, b) for a, b in zip(vals, freqs)).encode()
    p = tmp_path / "big.txt"
    p.write_bytes(text)
    v, f = api.read_histogram(p)
    rv, rf = ref_parse_histogram(ref, text)
    assert len(v) == n and same_bits(v, rv) and same_bits(f, rf)
    plan = api.Plan(v, f, 0.0)
    assert plan.n_bins == int((freqs > 0).sum()) and plan.n_cells == int(freqs.sum())


def _write_both(ref, tmp_path, values, freqs, ratios, tag):
    ours, theirs = tmp_path / ("ours_%s.txt" % tag), tmp_path / ("ref_%s.txt" % tag)
    api.write_histogram(str(ours), values, freqs, ratios)
    rv = np.ascontiguousarray(values, dtype=np.float64)
    rf = np.ascontiguousarray(np.maximum(freqs, 0), dtype=np.uint64)        # the product skips rows <= 0, as "> 0" does
    rr = None if ratios is None else np.ascontiguousarray(ratios, dtype=np.int64)
    rc = ref.ref_write_histogram(str(theirs).encode(), int(rr is not None), 0 if rr is None else rr.shape[1], len(rv),
                                 rv.ctypes.data_as(_f64p), rf.ctypes.data_as(_u64p),
                                 None if rr is None else rr.ctypes.data_as(_i64p))
    assert rc == 0
    return ours.read_bytes(), theirs.read_bytes()


def test_writer_bytes_equal_the_stream(ref, tmp_path):
    rng = np.random.default_rng(8)
    special = np.array([0.0, -0.0, 1.0, 0.5, 1e-5, 1e-4, 123456.7891234, 1234567890.0, 12345678901.0, 99999999995.0,
                        9999999999.5, 0.1, 1 / 3, 2 / 3, 1e10, 1e-10, 5e-324, 2.2250738585072014e-308,
                        1.7976931348623157e308, np.inf, -np.inf, 1009.0 / 2 ** 11, 9823.85, 8.144, 1e22, 1e23, 0.3])
    rand = rng.standard_normal(4000) * 10.0 ** rng.integers(-30, 31, 4000)
    halved = 10.0 ** rng.uniform(0, 4, 2000) / 2.0 ** rng.integers(0, 40, 2000)      # what real rows look like
    values = np.concatenate([special, rand, halved])
    freqs = rng.integers(-2, 2**62, len(values))
    freqs[rng.integers(0, len(values), 500)] = 0
    ours, theirs = _write_both(ref, tmp_path, values, freqs, None, "plain")
    assert ours == theirs and ours.count(b"\n") == int((freqs > 0).sum())
    for n_types in (1, 4, 64):
        ratios = rng.integers(-5, 2**62, (len(values), n_types))
        ours, theirs = _write_both(ref, tmp_path, values, freqs, ratios, "r%d" % n_types)
        assert ours == theirs
    # more than one buffer: 2e5 rows x 8 columns ~ 30 MB
    n = 200000
    values = 10.0 ** rng.uniform(-3, 5, n)
    freqs = rng.integers(0, 2**40, n)
    ratios = rng.integers(0, 2**40, (n, 8))
    ours, theirs = _write_both(ref, tmp_path, values, freqs, ratios, "big")
    assert len(ours) > (8 << 20) and ours == theirs


def test_written_rows_read_back(tmp_path):
    """output rows are a valid histogram input again (the reference's formats are symmetric for -r-less output)"""
    rng = np.random.default_rng(2)
    values = np.unique(np.round(10.0 ** rng.uniform(0, 4, 500), 6))
    freqs = rng.integers(1, 10**6, len(values))
    p = tmp_path / "o.txt"
    api.write_histogram(str(p), values, freqs)
    v, f = api.read_histogram(p)
    assert np.array_equal(f.astype(np.int64), freqs)
    assert np.allclose(v, values, rtol=1e-9, atol=0.0)       # 10 significant digits
    assert [float("%.10g" % x) for x in values] == v.tolist()


def test_reader_errors(tmp_path):
    with pytest.raises(api.ProcellError) as e:
        api.read_histogram(tmp_path / "missing.txt")
    assert e.value.code == _lib.ERR_IO and "cannot open histogram file" in str(e.value)
    with pytest.raises(api.ProcellError) as e:
        api.read_cell_types(tmp_path / "missing.txt")
    assert e.value.code == _lib.ERR_IO and "cannot open cell types file" in str(e.value)
    with pytest.raises(api.ProcellError) as e:
        api.write_histogram(str(tmp_path / "no_such_dir" / "o.txt"), [1.0], [1])
    assert e.value.code == _lib.ERR_IO
    with pytest.raises(api.ProcellError) as e:
        api.parse_cell_types(b"0.5 10 2\n0.4 -1 -1\n")
    assert e.value.code == _lib.ERR_PROPORTION
