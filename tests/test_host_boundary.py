"""The drop-in boundary on the host side (no GPU): C-ABI exports, text formats, plan, CLI argument handling."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cuda_pro_cell_b200 import _lib, api, synth

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "procell_b200.h").read_text()
    declared = set(re.findall(r"\b(procell_[a-z0-9_]+)\s*\(", header))
    lib = C.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), "libprocell_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "python binding and header disagree"
    assert b"sm_100a" in _lib.load().procell_version()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    v, f = synth.synthetic_histogram(100)
    plan = api.Plan(v, f, 1.0)
    with pytest.raises(api.ProcellError) as e:
        api.proliferate(plan, [synth.TYPES_CONFIG1], 10.0)
    assert e.value.code == _lib.ERR_CUDA and "no CPU path" in str(e.value)


def test_simulate_entry_point_checks_before_touching_a_device():
    """procell_simulate (SURVEY 8b's one-call entry): argument and proportion errors come back as codes, and on a box
    without a GPU a valid request fails loudly instead of falling back to a CPU path."""
    import torch
    v, f = synth.synthetic_histogram(100)
    with pytest.raises(api.ProcellError) as e:
        api.simulate(v, f, [(0.5, 10.0, 2.0), (0.4, -1.0, -1.0)], 10.0)
    assert e.value.code == _lib.ERR_PROPORTION and "does not sum to 1" in str(e.value)
    if not torch.cuda.is_available():
        with pytest.raises(api.ProcellError) as e:
            api.simulate(v, f, synth.TYPES_CONFIG1, 10.0)
        assert e.value.code == _lib.ERR_CUDA and "no CPU path" in str(e.value)


def test_histogram_reader_follows_operator_semantics(tmp_path):
    p = tmp_path / "h.txt"
    p.write_text("1.0 0\n8.144 53\n  9823.85\t274\n1e3 7 garbage 5\n12 3\n")
    v, f = api.read_histogram(p)
    # pairs are read until the first parse failure (parser.cu:103-106); zero-frequency lines are kept here
    assert v.tolist() == [1.0, 8.144, 9823.85, 1000.0] and f.tolist() == [0, 53, 274, 7]
    with pytest.raises(api.ProcellError):
        api.read_histogram(tmp_path / "missing.txt")


def test_cell_types_reader_and_proportion_check(tmp_path):
    p = tmp_path / "c.txt"
    p.write_text("0.53 48.33 21.6\n0.29 86.3 26.8\n0.18 -1 -1\n")
    t = api.read_cell_types(p)
    assert t.tolist() == [[0.53, 48.33, 21.6], [0.29, 86.3, 26.8], [0.18, -1.0, -1.0]]
    p.write_text("0.5 48.33 21.6\n0.3 86.3 26.8\n")
    with pytest.raises(api.ProcellError) as e:
        api.read_cell_types(p)
    assert e.value.code == _lib.ERR_PROPORTION
    assert "ERROR: proportion distribution of cell types does not sum to 1, aborting." in str(e.value)
    p.write_text("0.5 1 1\n0.499999995 2 2\n")       # |1 - sum| <= 1e-8 is accepted (parser.cu:60-61)
    assert len(api.read_cell_types(p)) == 2


def test_writer_format(tmp_path):
    out = tmp_path / "o.txt"
    rv = np.array([0.8360867925, 1.5, 1234567.891234, 1e-7, 12.0])
    rf = np.array([2, 0, 5, 1, 2**40], dtype=np.int64)
    rr = np.array([[2, 0], [0, 0], [1, 4], [0, 1], [2**40 - 1, 1]], dtype=np.int64)
    api.write_histogram(out, rv, rf, rr)
    # precision(10) default float format == %.10g, TAB separated, zero rows skipped, int64 counts (parser.cu:187-217)
    assert out.read_text() == ("0.8360867925\t2\t2\t0\n1234567.891\t5\t1\t4\n1e-07\t1\t0\t1\n"
                               "12\t1099511627776\t1099511627775\t1\n")
    api.write_histogram(out, rv, rf)
    assert out.read_text().splitlines()[0] == "0.8360867925\t2"


def test_plan_equals_oracle_plan(oracle):
    rng = np.random.default_rng(0)
    for n_lines, phi in ((1, 1.0), (50, 0.0), (300, 0.37), (1024, 1e-7), (40, 1e9)):
        v = np.round(10 ** rng.uniform(-2, 4, n_lines), 6)
        f = rng.integers(0, 50, n_lines).astype(np.uint64)
        a, b = api.Plan(v, f, phi), oracle.OraclePlan(v, f, phi)
        assert (a.n_bins, a.n_keys, a.n_rows, a.n_cells, a.phi) == (b.n_bins, b.n_keys, b.n_rows, b.n_cells, b.phi)
        assert np.array_equal(a.row_value, b.row_value) and np.array_equal(a.key_row, b.key_row)
        assert np.array_equal(a.bin_keybase, b.bin_keybase) and np.array_equal(a.bin_kdiv, b.bin_kdiv)
    # the bucketed key ordering (bins by mantissa, keys by exponent) and its fallback to a plain sort: values related by
    # powers of two (equal key values from different bins must share a row), unsorted and duplicate lines, a huge
    # dynamic range, subnormal / zero / negative values (fallback), halvings that run into the subnormal range
    cases = [
        ([8.0, 4.0, 2.0, 1.0, 3.0, 6.0, 12.0, 1.5], [1] * 8, 0.7),
        ([1000.0, 10.0, 1000.0, 500.0, 10.0, 2000.0], [3, 1, 2, 5, 7, 1], 3.0),
        ([1e300, 1e-300, 1.0, 3e150, 7e-200], [1, 2, 3, 4, 5], 1e-305),
        ([5e-324, 1e-310, 1.0, 4.0], [1, 1, 1, 1], 1e-320),
        ([0.0, -2.0, 8.0, 16.0], [1, 1, 1, 1], 1.0),
        ([3e-308, 6e-308, 1.2e-307], [2, 2, 2], 1e-323),
        (list(rng.uniform(1.0, 2.0, 64) * 2.0 ** rng.integers(-40, 40, 64)), [1] * 64, 1e-9),
        (list(np.repeat(rng.uniform(1.0, 2.0, 8), 8) * 2.0 ** np.tile(np.arange(8), 8)), [2] * 64, 0.01),
    ]
    for v, f, phi in cases:
        v, f = np.array(v, dtype=np.float64), np.array(f, dtype=np.uint64)
        a, b = api.Plan(v, f, phi), oracle.OraclePlan(v, f, phi)
        assert (a.n_bins, a.n_keys, a.n_rows, a.n_cells, a.phi) == (b.n_bins, b.n_keys, b.n_rows, b.n_cells, b.phi)
        assert np.array_equal(a.row_value, b.row_value) and np.array_equal(a.key_row, b.key_row)
        assert (np.diff(a.row_value) > 0).all()
    w = synth.workload(2)
    a, b = api.Plan(w.values, w.freqs, w.phi), oracle.OraclePlan(w.values, w.freqs, w.phi)
    assert np.array_equal(a.key_row, b.key_row) and a.n_keys == b.n_keys and np.array_equal(a.row_value, b.row_value)
    counts = rng.integers(0, 9, (a.n_keys, 4)).astype(np.int64)
    rf, rr = a.merge_rows(counts)
    assert int(rf.sum()) == int(counts.sum()) and np.array_equal(rr.sum(axis=1), rf)


def test_plan_limits():
    with pytest.raises(api.ProcellError):
        api.Plan(np.array([1.0]), np.array([2**33], dtype=np.uint64), 1.0)        # more than 2^32-1 seed cells
    p = api.Plan(np.array([1.0]), np.array([3], dtype=np.uint64), 1e-300)           # halvings capped at 63
    assert p.depth_capped and int(p.bin_kdiv[0]) == 63


def _cli(*args):
    return subprocess.run([str(_lib.CLI_PATH), *args], capture_output=True, text=True, timeout=60)


def test_lineage_depth_of_the_five_configs():
    """procell_plan_lineage_depth (no GPU needed): min(t_max / fastest mean, mean halvings phi allows) - what the library
    chooses the deep-tree kernel instance from (>= 6 and cells x 2^depth >= 5e8: configs 2 and 4; the choice itself is
    checked on the GPU, test_kernel_instance_chosen_per_workload)."""
    want = {1: (168.0 / 48.33, None), 2: (10.0, 10.0), 3: (None, None), 4: (30.0, 30.0)}
    got = {}
    for cfg in (1, 2, 3, 4):
        w = synth.workload(cfg, 1.0 if cfg != 3 else 0.01)
        plan = api.Plan(w.values, w.freqs, w.phi)
        got[cfg] = plan.lineage_depth(w.types, w.t_max)
        # independent restatement from the exported plan
        weights = np.diff(np.concatenate([[0], np.cumsum(w.freqs[w.freqs > 0])])).astype(np.float64)
        halvings = float(((plan.bin_kdiv & 63) * weights).sum() / weights.sum())
        fastest = min(m for _, m, _ in w.types[0] if m > 0)
        assert got[cfg] == pytest.approx(min(w.t_max / fastest, halvings), rel=1e-12)
    assert got[2] == want[2][0] and got[4] == want[4][0]            # time-bound: 240 / 24 and 720 / 24 generations
    assert 1.0 < got[1] < 2.0 and 1.5 < got[3] < 2.5                 # phi-bound: a lineage halves once or twice
    plan = api.Plan(np.array([8.0, 16.0]), np.array([3, 1], dtype=np.uint64), 0.5)
    assert plan.lineage_depth([(1.0, -1.0, -1.0)], 100.0) == 0.0     # nothing proliferates
    assert plan.lineage_depth([(1.0, 10.0, 1.0)], 0.0) == 0.0        # no time
    assert plan.lineage_depth([(0.5, 10.0, 1.0), (0.5, 5.0, 1.0)], 1e6) == pytest.approx((3 * 3 + 4 * 1) / 4.0)   # 8 / 2^(k+1) > 0.5: k = 0..2; 16: k = 0..3


def test_cli_messages_match_the_reference(tmp_path):
    """stdout + exit status 1, same strings as cmdargs.cpp:41-45,48-75,97-106 (README spellings accepted too)"""
    r = _cli("--bogus")
    assert (r.returncode, r.stdout) == (1, "Invalid option --bogus\n")
    r = _cli("-h", "a", "--histogram", "b")
    assert (r.returncode, r.stdout) == (1, "Option --histogram (-h) already given\n")
    r = _cli("-c")
    assert (r.returncode, r.stdout) == (1, "Option --cell-types (-c) requires a filename\n")
    r = _cli("-t", "-3")
    assert (r.returncode, r.stdout) == (1, "Option --time-max (-t) requires an integer value >= 0\n")
    r = _cli("-p", "0")
    assert (r.returncode, r.stdout) == (1, "Option --phi-min (-p) requires a double value > 0\n")
    r = _cli("-d", "24")
    assert (r.returncode, r.stdout) == (1, "Option --tree-depth (-d) requires an integer value >= 1 && <= 23\n")
    r = _cli("-r", "--track-ratio")
    assert (r.returncode, r.stdout) == (1, "Option --track-ratio (-r) already given\n")
    r = _cli("-o", "x")
    assert r.returncode == 1 and r.stdout.splitlines() == ["The following missing arguments are required:",
                                                           "--histogram (-h)", "--cell-types (-c)", "--t-max (-t)"]
    r = _cli("--help")
    assert r.returncode == 0 and "--histogram" in r.stdout and "--shard-level" in r.stdout
    r = _cli("--shard-level", "31")      # extension flags follow the reference's message style
    assert (r.returncode, r.stdout) == (1, "Option --shard-level requires an integer value >= 0 && <= 30\n")
    # -p is optional (README.md:98-102); both long spellings parse; bad proportions abort before any GPU work
    h, c = tmp_path / "h.txt", tmp_path / "c.txt"
    h.write_text("10 5\n")
    c.write_text("0.5 10 1\n0.3 20 2\n")
    r = _cli("--histogram", str(h), "--cell-types", str(c), "--time-max", "5", "--phi", "1", "--output", str(tmp_path / "o"),
             "-d", "7", "-r")
    assert (r.returncode, r.stdout) == (1, "ERROR: proportion distribution of cell types does not sum to 1, aborting.\n")
    r = _cli("-h", str(tmp_path / "nope"), "-c", str(c), "-t", "5")
    assert r.returncode == 1 and "cannot open histogram file" in r.stdout


_REF_BIN = ROOT / "oracle" / "_ref" / "procell_ref"

_CLI_CASES = [
    ["--bogus"], ["-x"], ["extra"], ["-t", "5", "extra"], ["--bogus", "--help"],
    ["-h"], ["-h", "a", "-h", "b"], ["--histogram", "a", "-h", "b"], ["-c"], ["-c", "a", "--cell-types", "b"],
    ["-t"], ["-t", "-3"], ["-t", "abc"], ["-t", "-0"], ["-t", "1e3x"], ["-t", "0", "-t", "1"], ["-t", "1.5", "--time-max", "2"],
    ["-p"], ["-p", "0"], ["-p", "-1"], ["-p", "abc"], ["-p", "1e-400"], ["-p", "1", "--phi-min", "2"],
    ["-d"], ["-d", "0"], ["-d", "24"], ["-d", "x"], ["-d", "3.7"], ["-d", "23", "-d", "2"],
    ["-r", "-r"], ["-r", "--track-ratio"], ["-o"], ["-o", "x", "-o", "y"], ["--output-histogram", "x", "-o", "y"],
    [], ["-o", "x"], ["-h", "a"], ["-h", "a", "-c", "b"], ["-p", "1"], ["-h", "a", "-c", "b", "-p", "1"],
    ["-h", "a", "-t", "5", "-p", "1"], ["-c", "b", "-t", "5", "-p", "1"],
]


@pytest.mark.skipif(not _REF_BIN.exists(), reason="the reference binary is built only where /root/reference exists (make -C oracle ref)")
@pytest.mark.parametrize("argv", _CLI_CASES, ids=lambda a: " ".join(a) or "(none)")
def test_cli_rejections_equal_the_reference_binary(argv):
    """Every command line the reference rejects BEFORE it touches CUDA (cmdargs.cpp:11-76 runs first, main.cu:18-24)
    is fed to the unmodified reference binary and to `procell`: same stdout, same exit status.  The one deliberate
    difference: -p is optional here, as the reference's README documents (README.md:98-102), so the reference's
    "--phi-min (-p)" line in the list of missing arguments has no counterpart."""
    ref = subprocess.run([str(_REF_BIN)] + argv, capture_output=True, text=True, timeout=60)
    new = _cli(*argv)
    want = "".join(l for l in ref.stdout.splitlines(keepends=True) if l != "--phi-min (-p)\n")
    if want == "The following missing arguments are required:\n":     # only -p was missing there: not an error here
        pytest.skip("the reference rejects this line for the missing -p alone")
    assert ref.returncode == 1 and ref.stderr == ""
    assert (new.returncode, new.stdout, new.stderr) == (1, want, "")


_TYPE_FILES = [          # (text of the cell-types file, does its proportion column sum to 1 within 1e-8 as read?)
    ("0.5 10 1\n0.5 20 2\n", True), ("0.5 10 1\n0.50000001 20 2\n", True), ("0.5 10 1\n0.500000011 20 2\n", False),
    ("0.5 10 1\n0.49999999 20 2\n", True), ("0.5 10 1\n0.499999989 20 2\n", False), ("0.1 1 1\n0.2 1 1\n0.7 -1 -1\n", True),
    ("1 10 1\n", True), ("", False), ("0.5 10\n", False), ("0.5 10 1 0.5 20 2", True), ("0.5 10 1\nx 1 1\n0.5 2 2\n", False),
    ("0.25 1 1\n0.25 1 1\n0.25 1 1\n0.25 1 1 trailing", True), ("1e0 5 5\n", True), ("0.3 1 1\n0.3 1 1\n0.4\n", False),
]


@pytest.mark.skipif(not _REF_BIN.exists(), reason="the reference binary is built only where /root/reference exists (make -C oracle ref)")
@pytest.mark.parametrize("text,sums", _TYPE_FILES)
def test_proportion_check_equals_the_reference_binary(tmp_path, text, sums):
    """The reference reads both files and checks the proportions (parser.cu:46-66,156-185) before its first CUDA call,
    so on a box without a GPU its verdict on a cell-types file is observable: rejected with the message on stdout and
    status 1, or accepted (it then dies in its first CUDA call).  `procell` must draw the same line at 1e-8."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("on a GPU box the reference would run the simulation")
    h, c = tmp_path / "h.txt", tmp_path / "c.txt"
    h.write_text("10 5\n")
    c.write_text(text)
    argv = ["-h", str(h), "-c", str(c), "-t", "5", "-p", "1"]
    ref = subprocess.run([str(_REF_BIN)] + argv, capture_output=True, text=True, timeout=60)
    new = _cli(*argv)
    msg = "ERROR: proportion distribution of cell types does not sum to 1, aborting.\n"
    assert ((ref.returncode, ref.stdout) == (1, msg)) == (not sums)
    if sums:
        assert new.returncode == 1 and "no CUDA device" in new.stdout and msg not in new.stdout
    else:
        assert (new.returncode, new.stdout) == (1, msg)


def test_python_structs_match_the_c_header(tmp_path):
    """ctypes mirrors of the C-ABI structs against the header itself: a C program prints sizeof / offsetof of every
    struct that crosses the boundary by pointer; a field added on one side only would shift everything behind it."""
    import shutil
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"procell_cell_type": _lib.CellType, "procell_sim_params": _lib.SimParams, "procell_run_stats": _lib.RunStats,
               "procell_input": _lib.SimInput, "procell_output": _lib.SimOutput}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "procell_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["return 0;", "}"]
    src, exe = tmp_path / "abi.c", tmp_path / "abi"
    src.write_text("\n".join(lines))
    subprocess.run([cc, "-std=c11", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, "%s.%s" % (cname, fname)
    # and the header has no field the mirror lacks: same number of members (counted from the header text)
    header = (ROOT / "include" / "procell_b200.h").read_text()
    for cname, cls in structs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        members = [d for decl in body.split(";") if decl.strip() for d in decl.split(",")]
        assert len(members) == len(cls._fields_), "%s: header has %d members, the ctypes mirror %d" % (cname, len(members), len(cls._fields_))


@pytest.mark.skipif(not _REF_BIN.exists(), reason="the reference binary is built only where /root/reference exists (make -C oracle ref)")
def test_random_command_lines_are_rejected_like_the_reference_binary():
    """300 random sequences of the reference's flags, stray words and values: whatever the reference rejects (status 1,
    before any CUDA call) `procell` rejects with the same stdout - same precedence between "invalid option", "already
    given", "requires ..." and the list of missing arguments.  (-p is optional here: that line is dropped.)"""
    import random
    rng = random.Random(7)
    flags = ["-h", "--histogram", "-c", "--cell-types", "-t", "--time-max", "-o", "--output-histogram", "-p", "--phi-min",
             "-d", "--tree-depth", "-r", "--track-ratio", "--bogus", "-x", "extra"]
    vals = ["a", "b", "1", "0", "-1", "1.5", "abc", "24", "23", "0.0", "1e-3", "", "-r", "-t", "3.7", "1e400", "-0"]
    compared = 0
    for _ in range(300):
        argv = []
        for _ in range(rng.randrange(0, 7)):
            argv.append(rng.choice(flags))
            if rng.random() < 0.7:
                argv.append(rng.choice(vals))
        ref = subprocess.run([str(_REF_BIN)] + argv, capture_output=True, text=True, timeout=60)
        if ref.returncode != 1:
            continue                    # accepted: the reference goes on to CUDA and dies there without a GPU
        want = "".join(l for l in ref.stdout.splitlines(keepends=True) if l != "--phi-min (-p)\n")
        if want == "The following missing arguments are required:\n":
            continue
        new = _cli(*argv)
        assert (new.returncode, new.stdout) == (1, want), argv
        compared += 1
    assert compared > 200
