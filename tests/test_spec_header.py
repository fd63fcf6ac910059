"""The arithmetic header the kernels include (cuda_pro_cell_b200/csrc/procell_spec.h), built for the HOST by a
test-only shim, must agree bit-for-bit with the oracle's independent restatement.  Catches a spec typo without a GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


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
    return S


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
