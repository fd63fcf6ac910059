#!/usr/bin/env python3
"""Host-side throughput of the streaming text I/O (csrc/textio.cpp) next to the iostream restatement of the
reference's loops (tests/textio_ref.cpp), on the largest inputs the key layout admits.  No GPU.

  python tools/textio_bench.py > profiles/r1_textio_host.json
"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cuda_pro_cell_b200 import api  # noqa: E402

_f64p, _u64p, _i64p = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)


def build_ref():
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    so = out / "libtextio_ref.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(ROOT / "tests" / "textio_ref.cpp")], check=True)
    R = C.CDLL(str(so))
    R.ref_read_histogram.restype = C.c_size_t
    R.ref_read_histogram.argtypes = [C.c_char_p, _f64p, _u64p, C.c_size_t]
    R.ref_write_histogram.argtypes = [C.c_char_p, C.c_int, C.c_int32, C.c_size_t, _f64p, _u64p, _i64p]
    return R


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    R = build_ref()
    rng = np.random.default_rng(1)
    tmp = Path(tempfile.mkdtemp(prefix="textio_"))
    res = {"what": "host text I/O, best of 5, page cache warm", "cpu": os.cpu_count()}

    n = 65535                                   # the most histogram lines the key layout admits
    vals = np.sort(rng.uniform(1.0, 1e5, n))
    freqs = rng.integers(1, 2**31, n)
    h = tmp / "h.txt"
    h.write_text("".join("%.10g %d\n" % (a, b) for a, b in zip(vals, freqs)))
    size = h.stat().st_size
    rv, rf = np.zeros(n), np.zeros(n, dtype=np.uint64)
    t_ours = best(lambda: api.read_histogram(h))
    t_ref = best(lambda: R.ref_read_histogram(str(h).encode(), rv.ctypes.data_as(_f64p), rf.ctypes.data_as(_u64p), n))
    res["read_histogram"] = {"lines": n, "bytes": size, "ours_ms": 1e3 * t_ours, "iostream_ms": 1e3 * t_ref,
                             "ours_MBps": size / t_ours / 1e6, "iostream_MBps": size / t_ref / 1e6,
                             "speedup": t_ref / t_ours}

    for rows, n_types in ((4420, 4), (1_000_000, 8), (2_000_000, 64)):
        values = np.sort(10.0 ** rng.uniform(-3, 5, rows))
        fr = rng.integers(1, 2**40, rows)
        ra = rng.integers(0, 2**40, (rows, n_types))
        o1, o2 = tmp / "o1.txt", tmp / "o2.txt"
        reps = 5 if rows <= 1_000_000 else 2
        t_ours = best(lambda: api.write_histogram(str(o1), values, fr, ra), reps)
        fu = fr.astype(np.uint64)
        t_ref = best(lambda: R.ref_write_histogram(str(o2).encode(), 1, n_types, rows, values.ctypes.data_as(_f64p),
                                                   fu.ctypes.data_as(_u64p), ra.ctypes.data_as(_i64p)), reps)
        assert o1.read_bytes() == o2.read_bytes()
        size = o1.stat().st_size
        res["write_rows_%d_types_%d" % (rows, n_types)] = {
            "rows": rows, "n_types": n_types, "bytes": size, "ours_ms": 1e3 * t_ours, "iostream_ms": 1e3 * t_ref,
            "ours_MBps": size / t_ours / 1e6, "iostream_MBps": size / t_ref / 1e6, "speedup": t_ref / t_ours}
        o1.unlink(); o2.unlink()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
