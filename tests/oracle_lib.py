"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure only: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


def build_oracle(target: str = "liboracle.so") -> Path:
    so = ORACLE_DIR / target
    srcs = [ORACLE_DIR / "procell_oracle.c", ORACLE_DIR / "xorwow_ref.c", ORACLE_DIR / "oracle_math_tables.inc"]
    newest = max((s.stat().st_mtime for s in srcs if s.exists()), default=0)
    if not so.exists() or so.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(ORACLE_DIR), target], check=True, capture_output=True)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build_oracle()))
        L.oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.oracle_uniform53.argtypes = [C.c_uint32, C.c_uint32]
        L.oracle_uniform53.restype = C.c_double
        L.oracle_neg2log.argtypes = [C.c_double]
        L.oracle_neg2log.restype = C.c_double
        L.oracle_sincos2pi.argtypes = [C.c_uint64, _f64p, _f64p]
        L.oracle_normal_pair.argtypes = [_u32p, C.c_double, _f64p]
        L.oracle_zig_trial.argtypes = [_u32p, C.c_uint, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, _f64p]
        L.oracle_zig_trial.restype = C.c_int
        L.oracle_zig_fill.argtypes = [C.c_uint64, C.c_uint64, _f64p, _u64p]
        L.oracle_plan_create.argtypes = [_f64p, _u64p, C.c_size_t, C.c_double]
        L.oracle_plan_create.restype = C.c_void_p
        L.oracle_plan_free.argtypes = [C.c_void_p]
        for name, rt in (("n_bins", C.c_size_t), ("n_keys", C.c_size_t), ("n_rows", C.c_size_t),
                         ("n_cells", C.c_uint64), ("phi", C.c_double), ("depth_capped", C.c_int)):
            f = getattr(L, "oracle_plan_" + name)
            f.argtypes = [C.c_void_p]
            f.restype = rt
        L.oracle_plan_export.argtypes = [C.c_void_p, _f64p, _u32p, _u32p, _u8p]
        L.oracle_check_proportions.argtypes = [_f64p, C.c_size_t]
        L.oracle_check_proportions.restype = C.c_int
        L.oracle_simulate.argtypes = [C.c_void_p, _f64p, C.c_size_t, C.c_size_t, C.c_double, C.c_uint64, C.c_int,
                                      C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                      _i64p, _i64p]
        L.oracle_simulate.restype = C.c_int
        L.oracle_merge_rows.argtypes = [C.c_void_p, _i64p, C.c_size_t, _i64p, _i64p]
        _lib = L
    return _lib


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return tuple(o)


def sincos2pi(v: int):
    s, c = C.c_double(), C.c_double()
    lib().oracle_sincos2pi(C.c_uint64(v), C.byref(s), C.byref(c))
    return s.value, c.value


def normal_pair(words, u_forced=0.0):
    w = (C.c_uint32 * 4)(*words)
    z = (C.c_double * 2)()
    lib().oracle_normal_pair(w, u_forced, z)
    return z[0], z[1]


def zig_trial(words, c, root=0, set_=0, retry=0, heap=1, seed=0):
    """one whole ziggurat trial for daughter c: (accepted, z)"""
    w = (C.c_uint32 * 4)(*words)
    z = C.c_double(0.0)
    ok = lib().oracle_zig_trial(w, c, root, set_, retry, C.c_uint64(heap), C.c_uint64(seed), C.byref(z))
    return bool(ok), z.value


def zig_fill(seed: int, n: int):
    """n standard normals as divisions draw them (redraw on rejection) and the number of trials made"""
    out = np.zeros(n, dtype=np.float64)
    trials = (C.c_uint64 * 1)()
    lib().oracle_zig_fill(C.c_uint64(seed), C.c_uint64(n), out.ctypes.data_as(_f64p), trials)
    return out, int(trials[0])


class OraclePlan:
    def __init__(self, values, freqs, phi=0.0):
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.freqs = np.ascontiguousarray(freqs, dtype=np.uint64)
        L = lib()
        self.h = L.oracle_plan_create(self.values.ctypes.data_as(_f64p), self.freqs.ctypes.data_as(_u64p),
                                      len(self.values), float(phi))
        self.n_bins = L.oracle_plan_n_bins(self.h)
        self.n_keys = L.oracle_plan_n_keys(self.h)
        self.n_rows = L.oracle_plan_n_rows(self.h)
        self.n_cells = L.oracle_plan_n_cells(self.h)
        self.phi = L.oracle_plan_phi(self.h)
        self.depth_capped = bool(L.oracle_plan_depth_capped(self.h))
        self.row_value = np.zeros(self.n_rows, dtype=np.float64)
        self.key_row = np.zeros(self.n_keys, dtype=np.uint32)
        self.bin_keybase = np.zeros(self.n_bins, dtype=np.uint32)
        self.bin_kdiv = np.zeros(self.n_bins, dtype=np.uint8)
        L.oracle_plan_export(self.h, self.row_value.ctypes.data_as(_f64p), self.key_row.ctypes.data_as(_u32p),
                             self.bin_keybase.ctypes.data_as(_u32p), self.bin_kdiv.ctypes.data_as(_u8p))

    def __del__(self):
        try:
            lib().oracle_plan_free(self.h)
        except Exception:
            pass


def simulate(plan: OraclePlan, types, t_max, seed, refcompat=False, root_begin=0, root_end=None,
             shard=(0, 1, 1), n_threads=0, shard_level=0):
    """types: array [n_sets][n_types][3] or [n_types][3]; shard = (rank, world, unit); shard_level >= 1 selects subtree
    sharding (every rank walks the first levels of every lineage, see oracle_simulate).  Returns dict(counts[S,K,T], divisions[S],
    row_freq[S,R], row_ratio[S,R,T])."""
    t = np.ascontiguousarray(types, dtype=np.float64)
    if t.ndim == 2:
        t = t[None]
    n_sets, n_types, _ = t.shape
    counts = np.zeros((n_sets, plan.n_keys, n_types), dtype=np.int64)
    divisions = np.zeros(n_sets, dtype=np.int64)
    if root_end is None:
        root_end = plan.n_cells
    rc = lib().oracle_simulate(plan.h, t.ctypes.data_as(_f64p), n_types, n_sets, float(t_max), C.c_uint64(seed),
                               int(bool(refcompat)), C.c_uint64(root_begin), C.c_uint64(root_end),
                               int(shard[2]) or 1, int(shard[1]) or 1, int(shard[0]), int(shard_level), int(n_threads),
                               counts.ctypes.data_as(_i64p), divisions.ctypes.data_as(_i64p))
    if rc != 0:
        raise RuntimeError("oracle_simulate failed rc=%d" % rc)
    row_freq = np.zeros((n_sets, plan.n_rows), dtype=np.int64)
    row_ratio = np.zeros((n_sets, plan.n_rows, n_types), dtype=np.int64)
    for s in range(n_sets):
        lib().oracle_merge_rows(plan.h, counts[s].ctypes.data_as(_i64p), n_types,
                                row_freq[s].ctypes.data_as(_i64p), row_ratio[s].ctypes.data_as(_i64p))
    return dict(counts=counts, divisions=divisions, row_freq=row_freq, row_ratio=row_ratio)


def n_host_threads() -> int:
    return len(os.sched_getaffinity(0))


# ---- the reference's own XORWOW stream (oracle/xorwow_ref.c), pinned by tests/golden/ref_*.json -------------------
_xlib = None


def xlib():
    global _xlib
    if _xlib is None:
        X = C.CDLL(str(build_oracle("libxorwow_ref.so")))
        _i32p = C.POINTER(C.c_int32)
        X.xorwow_ref_simulate.restype = C.c_long
        X.xorwow_ref_simulate.argtypes = [_f64p, _u64p, C.c_size_t, _f64p, C.c_size_t, C.c_double, C.c_double,
                                          C.c_uint64, C.c_uint64, C.c_int, C.c_size_t, _f64p, _u64p, _i32p, _u64p]
        X.xorwow_uniform.restype = C.c_double
        X.xorwow_uniform.argtypes = [C.c_uint64]
        X.xorwow_normal.restype = C.c_double
        X.xorwow_normal.argtypes = [C.c_uint64, C.c_double, C.c_double]
        _xlib = X
    return _xlib


def xorwow_simulate(values, freqs, types, t_max, phi, T0, T1=None, max_depth=23):
    """The reference as written, for wall-clock seeds T0 (population) and T1 (iteration).  Returns dict(row_value,
    row_freq, row_ratio, divisions) with zero rows included, or None if a cell outlives max_depth levels."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    t = np.ascontiguousarray(types, dtype=np.float64)
    cap = 1
    for val, fr in zip(v, f):
        c = float(val)
        while fr > 0 and c >= phi and cap < (1 << 24):
            cap += 1
            c /= 2
    rv = np.zeros(cap, dtype=np.float64)
    rf = np.zeros(cap, dtype=np.uint64)
    rr = np.zeros((cap, len(t)), dtype=np.int32)
    dv = C.c_uint64()
    n = xlib().xorwow_ref_simulate(v.ctypes.data_as(_f64p), f.ctypes.data_as(_u64p), len(v), t.ctypes.data_as(_f64p),
                                   len(t), float(t_max), float(phi), int(T0), int(T0 if T1 is None else T1),
                                   int(max_depth), cap, rv.ctypes.data_as(_f64p), rf.ctypes.data_as(_u64p),
                                   rr.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(dv))
    if n == -2:
        return None
    if n < 0:
        raise RuntimeError("xorwow_ref_simulate failed rc=%d" % n)
    return dict(row_value=rv[:n], row_freq=rf[:n].astype(np.int64), row_ratio=rr[:n].astype(np.int64),
                divisions=int(dv.value))
