"""The arithmetic header the kernels include (cuda_pro_cell_b200/csrc/procell_spec.h), built for the HOST by a
test-only shim, must agree bit-for-bit with the oracle's independent restatement.  Catches a spec typo without a GPU."""
import ctypes as C
import re
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
_INC = (ROOT / "cuda_pro_cell_b200" / "csrc" / "procell_math_tables.inc").read_text()
ZIG_BITS = int(re.search(r"#define PCM_ZIG_N_BITS\s+(\d+)", _INC).group(1))        # layers of the ziggurat = 2^ZIG_BITS
ZIG_N = 1 << ZIG_BITS
ZIG_R = struct.unpack("<d", struct.pack("<Q", int(re.search(r"#define PCM_BITS_ZIG_R\s+0x([0-9A-F]+)ULL", _INC).group(1), 16)))[0]


@pytest.fixture(scope="module")
def shim():
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    so = out / "libspec_shim.so"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", str(so),
                    str(ROOT / "tests" / "spec_shim.cpp")], check=True)
    S = C.CDLL(str(so))
    S.shim_u53.restype = C.c_double
    S.shim_u53.argtypes = [C.c_uint32, C.c_uint32]
    S.shim_neg2log.restype = C.c_double
    S.shim_neg2log.argtypes = [C.c_double]
    S.shim_timer.restype = C.c_double
    S.shim_timer.argtypes = [C.c_double] * 3
    S.shim_draw.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
    S.shim_u32unit.restype = C.c_double
    S.shim_u32unit.argtypes = [C.c_uint32]
    S.shim_seed_normal.restype = C.c_double
    S.shim_seed_normal.argtypes = [C.POINTER(C.c_uint32), C.c_double]
    S.shim_zig_trial.restype = C.c_int
    S.shim_zig_trial.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                 C.POINTER(C.c_double)]
    S.shim_zig_fast.restype = C.c_int
    S.shim_zig_fast.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    S.shim_zig_fast_kernel_form.restype = C.c_int
    S.shim_zig_fast_kernel_form.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    S.shim_zig_fast_forms_differ.restype = C.c_uint64
    S.shim_zig_fast_forms_differ.argtypes = [C.c_uint64, C.c_uint64]
    return S


def test_kernel_form_of_the_fast_ziggurat_test_is_the_spec_bitwise(shim):
    """The kernels put the draw's sign on the multiplicand and test |x| < x_next (sim_kernels.cu: zig_fast_smem) where
    the spec ORs the sign onto the product (procell_spec.h: pcs_zig_fast).  (-M) * x_i = -(M * x_i) exactly, so both the
    verdict and the bits of z - a signed zero included - must agree: 2e7 random draws and the edge mantissas (0, 1, the
    subnormal / normal boundary, all ones) of every layer with either sign."""
    assert shim.shim_zig_fast_forms_differ(20_000_000, 0x9E3779B97F4A7C15) == 0
    z, zk = C.c_double(1.0), C.c_double(1.0)
    hi = 1 << 31                                              # layer 0, mantissa 0, negative: z = -0.0 in both forms
    assert shim.shim_zig_fast(0, hi, C.byref(z)) == shim.shim_zig_fast_kernel_form(0, hi, C.byref(zk)) == 1
    assert struct.pack("<d", z.value) == struct.pack("<d", zk.value) == struct.pack("<d", -0.0)


def test_seed_cell_draws_equal_the_oracle_bitwise(shim, oracle):
    """the seed cell's ONE block: 32-bit uniforms (type, age, radius) and the first timer's normal from words z, w"""
    rng = np.random.default_rng(5)
    L = oracle.lib()
    L.oracle_uniform32.restype = C.c_double
    L.oracle_uniform32.argtypes = [C.c_uint32]
    for m in [0, 1, 2**31, 2**32 - 1] + [int(x) for x in rng.integers(0, 2**32, 5000)]:
        u = shim.shim_u32unit(m)
        assert u == L.oracle_uniform32(m) == (2 * m + 1) / 2.0**33 and 0.0 < u < 1.0
    for _ in range(5000):
        w = [int(x) for x in rng.integers(0, 2**32, 4)]
        for forced in (0.0, 0.41):
            got = shim.shim_seed_normal((C.c_uint32 * 4)(*w), C.c_double(forced))
            u = forced if forced > 0 else L.oracle_uniform32(w[2])
            s, c = oracle.sincos2pi(w[3] << 32)
            assert got == np.sqrt(L.oracle_neg2log(u)) * c


def test_integer_type_thresholds_are_exactly_the_double_compare(shim):
    """The kernels pick the seed cell's type by comparing the 32-bit type word x with integer thresholds
    (procell_type_threshold) instead of comparing the uniform (2x + 1) / 2^33 with the running proportion sums as
    cell.cu:81-104 does: "u(x) < cum" must hold for exactly the x below the threshold - checked at the boundary for
    proportions a user would write, for sums a hair around 1, for tiny and awkward values."""
    from cuda_pro_cell_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    cums = [0.53, 0.82, 1.0, 1.0 - 1e-9, 1.0 + 1e-9, 0.4, 0.65, 0.17, 1 / 3, 2 / 3, 0.5, 0.25, 2.0**-33, 2.0**-32, 3 * 2.0**-34,
            2.0**-40, 1e-300, 0.9999999999, float(np.nextafter(1.0, 0.0)), float(np.nextafter(0.5, 1.0))]
    cums += [float(x) for x in rng.random(3000)] + [float(x) * 2.0**-20 for x in rng.random(200)]
    cums += [(2 * int(x) + 1) / 2.0**33 for x in rng.integers(0, 2**32, 500)]          # a sum that IS a uniform: strict "<"
    for c in cums:
        thr = lib.procell_type_threshold(c)
        assert 0 <= thr <= 2**32
        if thr > 0:
            assert shim.shim_u32unit(thr - 1) < c          # the last x below the sum
        if thr < 2**32:
            assert not (shim.shim_u32unit(thr) < c)        # the first x that is not
    assert lib.procell_type_threshold(0.0) == 0 and lib.procell_type_threshold(-1.0) == 0
    assert lib.procell_type_threshold(5.0) == 2**32


def test_header_equals_oracle_bitwise(shim, oracle):
    rng = np.random.default_rng(2)
    L = oracle.lib()
    for _ in range(20000):
        w = [int(x) for x in rng.integers(0, 2**32, 4)]
        z = (C.c_double * 2)()
        shim.shim_normal_pair((C.c_uint32 * 4)(*w), C.c_double(0.0), z)
        assert (z[0], z[1]) == oracle.normal_pair(w)
        u = L.oracle_uniform53(w[0], w[1])
        assert shim.shim_u53(w[0], w[1]) == u
        assert shim.shim_neg2log(u) == L.oracle_neg2log(u)
        zf = (C.c_double * 2)()
        shim.shim_normal_pair((C.c_uint32 * 4)(*w), C.c_double(0.37), zf)
        assert (zf[0], zf[1]) == oracle.normal_pair(w, 0.37)


def test_ziggurat_trials_equal_the_oracle_bitwise(shim, oracle):
    """Division timers: the header's ziggurat trial (fast test, wedge test, tail sampler) against the oracle's independent
    restatement, bit for bit - on random blocks (99 % fast), on blocks forced onto the edge of a layer (wedge tests,
    both outcomes), on the base strip beyond r (the tail sampler) and on the top layer (always a wedge test)."""
    rng = np.random.default_rng(9)
    seen = {"fast": 0, "wedge_acc": 0, "wedge_rej": 0, "tail": 0}

    def check(w, c, root, st, retry, heap, seed):
        z = C.c_double(0.0)
        got = shim.shim_zig_trial((C.c_uint32 * 4)(*w), c, root, st, retry, C.c_uint64(heap), C.c_uint64(seed), C.byref(z))
        ok, want = oracle.zig_trial(w, c, root, st, retry, heap, seed)
        assert bool(got) == ok
        if ok:
            assert z.value == want and np.isfinite(want)
        zf = C.c_double(0.0)
        fast = shim.shim_zig_fast(w[2 * c], w[2 * c + 1], C.byref(zf))
        layer = (w[2 * c + 1] >> (31 - ZIG_BITS)) & (ZIG_N - 1)
        if fast:
            assert ok and zf.value == want
            seen["fast"] += 1
        elif layer == 0:
            assert ok and abs(want) >= ZIG_R          # the tail sampler always accepts, beyond r
            assert (want < 0) == bool(w[2 * c + 1] >> 31)
            seen["tail"] += 1
        else:
            if ok:
                assert want == zf.value                            # a wedge test does not move the point
            seen["wedge_acc" if ok else "wedge_rej"] += 1

    for _ in range(20000):
        w = [int(x) for x in rng.integers(0, 2**32, 4)]
        check(w, int(rng.integers(0, 2)), int(rng.integers(0, 2**32)), int(rng.integers(0, 65536)), int(rng.integers(0, 255)),
              int(rng.integers(1, 2**62)), int(rng.integers(0, 2**63)))
    for _ in range(20000):          # mantissa close to 1: the point is near the outer edge of its layer
        w = [int(x) for x in rng.integers(0, 2**32, 4)]
        c = int(rng.integers(0, 2))
        layer = int(rng.choice([0, 0, 1, 2, ZIG_N // 2 - 1, ZIG_N - 2, ZIG_N - 1, int(rng.integers(0, ZIG_N))]))
        frac = 0x1FFFFF - int(rng.integers(0, 0x60000))
        w[2 * c + 1] = (int(rng.integers(0, 2)) << 31) | (layer << (31 - ZIG_BITS)) | (int(rng.integers(0, 1 << (10 - ZIG_BITS))) << 21) | frac
        check(w, c, int(rng.integers(0, 2**32)), int(rng.integers(0, 65536)), int(rng.integers(0, 255)),
              int(rng.integers(1, 2**62)), int(rng.integers(0, 2**63)))
    assert seen["fast"] > 19000 and seen["wedge_acc"] > 1000 and seen["wedge_rej"] > 1000 and seen["tail"] > 1000, seen


def test_header_counter_layout(shim, oracle):
    """ctr = {root, set | retry<<16 | tag<<24, heap_lo, heap_hi}, key = {seed_lo, seed_hi}"""
    rng = np.random.default_rng(4)
    for _ in range(500):
        root = int(rng.integers(0, 2**32))
        st, retry, tag = int(rng.integers(0, 65536)), int(rng.integers(0, 256)), int(rng.integers(0, 2))
        heap, seed = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**63))
        out = (C.c_uint32 * 4)()
        shim.shim_draw(root, st, retry, tag, C.c_uint64(heap), C.c_uint64(seed), out)
        want = oracle.philox((root, st | (retry << 16) | (tag << 24), heap & 0xFFFFFFFF, heap >> 32),
                             (seed & 0xFFFFFFFF, seed >> 32))
        assert tuple(out) == want
