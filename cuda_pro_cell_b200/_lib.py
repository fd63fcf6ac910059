"""ctypes binding of libprocell_b200.so (the C ABI declared in include/procell_b200.h)."""
import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
import os as _os

# PROCELL_LIB selects an alternative in-tree build of the same library (A/B experiments); default: the product build
LIB_PATH = PKG_DIR / _os.environ.get("PROCELL_LIB", "libprocell_b200.so")
CLI_PATH = PKG_DIR / "procell"

OK, ERR_ARG, ERR_CUDA, ERR_PROPORTION, ERR_IO, ERR_OVERFLOW = 0, -1, -2, -3, -4, -5
SEEDING_IDEAL, SEEDING_REFCOMPAT = 0, 1
KERNEL_COOP, KERNEL_SIMPLE = 0, 1


class CellType(C.Structure):
    _fields_ = [("proportion", C.c_double), ("mean", C.c_double), ("stddev", C.c_double)]


class SimParams(C.Structure):
    _fields_ = [("types", C.POINTER(CellType)), ("n_types", C.c_size_t), ("n_sets", C.c_size_t),
                ("t_max", C.c_double), ("seed", C.c_uint64), ("seeding_mode", C.c_int), ("kernel", C.c_int),
                ("shard_rank", C.c_uint32), ("shard_world", C.c_uint32), ("shard_unit", C.c_uint32),
                ("checkpoints", C.POINTER(C.c_double)), ("n_checkpoints", C.c_size_t), ("shard_level", C.c_uint32)]


class RunStats(C.Structure):
    _fields_ = [("divisions", C.c_int64), ("kernel_ms", C.c_double), ("n_launches", C.c_int),
                ("grid", C.c_int), ("block", C.c_int), ("smem_bytes", C.c_int), ("donations", C.c_int64),
                ("seed_phase_us", C.c_double), ("total_us", C.c_double), ("idle_warp_us", C.c_double),
                ("idle_waits", C.c_int64)]


class SimInput(C.Structure):
    _fields_ = [("bin_value", C.POINTER(C.c_double)), ("bin_freq", C.POINTER(C.c_uint64)), ("n_bins", C.c_size_t),
                ("types", C.POINTER(CellType)), ("n_types", C.c_size_t), ("n_param_sets", C.c_size_t),
                ("t_max", C.c_double), ("phi", C.c_double), ("track_ratio", C.c_int), ("seed", C.c_uint64),
                ("n_gpus", C.c_int), ("seeding_mode", C.c_int)]


class SimOutput(C.Structure):
    _fields_ = [("value", C.POINTER(C.c_double)), ("freq", C.POINTER(C.c_int64)), ("ratio", C.POINTER(C.c_int64)),
                ("n_rows", C.c_size_t), ("divisions", C.c_int64), ("kernel_ms", C.c_double)]


_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)

# every symbol include/procell_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "procell_last_error": (C.c_char_p, []),
    "procell_version": (C.c_char_p, []),
    "procell_read_histogram": (C.c_int, [C.c_char_p, C.POINTER(_f64p), C.POINTER(_u64p), C.POINTER(C.c_size_t)]),
    "procell_read_cell_types": (C.c_int, [C.c_char_p, C.POINTER(C.POINTER(CellType)), C.POINTER(C.c_size_t)]),
    "procell_parse_histogram": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_f64p), C.POINTER(_u64p), C.POINTER(C.c_size_t)]),
    "procell_parse_cell_types": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(CellType)), C.POINTER(C.c_size_t)]),
    "procell_write_histogram": (C.c_int, [C.c_char_p, C.c_int, C.c_size_t, C.c_size_t, _f64p, _i64p, _i64p]),
    "procell_free": (None, [C.c_void_p]),
    "procell_type_threshold": (C.c_uint64, [C.c_double]),
    "procell_check_proportions": (C.c_int, [C.POINTER(CellType), C.c_size_t]),
    "procell_plan_create": (C.c_int, [_f64p, _u64p, C.c_size_t, C.c_double, C.POINTER(C.c_void_p)]),
    "procell_plan_destroy": (None, [C.c_void_p]),
    "procell_plan_n_bins": (C.c_size_t, [C.c_void_p]),
    "procell_plan_n_keys": (C.c_size_t, [C.c_void_p]),
    "procell_plan_n_rows": (C.c_size_t, [C.c_void_p]),
    "procell_plan_n_cells": (C.c_uint64, [C.c_void_p]),
    "procell_plan_phi": (C.c_double, [C.c_void_p]),
    "procell_plan_depth_capped": (C.c_int, [C.c_void_p]),
    "procell_plan_lineage_depth": (C.c_double, [C.c_void_p, _f64p, C.c_size_t, C.c_double]),
    "procell_plan_export": (C.c_int, [C.c_void_p, _f64p, _u32p, _u32p, _u8p]),
    "procell_merge_rows": (C.c_int, [C.c_void_p, _i64p, C.c_size_t, _i64p, _i64p]),
    "procell_proliferate": (C.c_int, [C.c_void_p, C.POINTER(SimParams), C.c_int, _i64p, _i64p, C.POINTER(RunStats)]),
    "procell_proliferate_multi": (C.c_int, [C.c_void_p, C.POINTER(SimParams), C.c_int, _i64p, _i64p, C.POINTER(RunStats)]),
    "procell_simulate": (C.c_int, [C.POINTER(SimInput), C.POINTER(SimOutput)]),
    "procell_output_free": (None, [C.POINTER(SimOutput)]),
    "procell_engine_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "procell_engine_destroy": (None, [C.c_void_p]),
    "procell_engine_load": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SimParams)]),
    "procell_engine_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "procell_engine_finish": (C.c_int, [C.c_void_p, C.c_void_p, _i64p, _i64p, C.POINTER(RunStats)]),
    "procell_engine_counts_len": (C.c_size_t, [C.c_void_p]),
    "procell_engine_set_target": (C.c_int, [C.c_void_p, _f64p, _u64p, C.c_size_t]),
    "procell_engine_fitness": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _f64p]),
    "procell_engine_fitness_in_launch": (C.c_int, [C.c_void_p]),
    "procell_engine_kernel_mode": (C.c_int, [C.c_void_p]),
    "procell_rng_ceiling_variants": (C.c_int, [C.c_int, C.c_int, _f64p, _f64p]),
    "procell_rng_ceiling": (C.c_int, [C.c_int, C.c_int, _f64p, _f64p]),
    "procell_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
}

_lib = None


class ProcellError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("procell error %d: %s" % (code, message))
        self.code = code


def load():
    """Load the shared library; raises (never falls back) when it is missing or lacks a symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError("%s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                          "`make -C cuda_pro_cell_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise ProcellError(rc, load().procell_last_error().decode("utf-8", "replace"))
