"""Host-side mirror of the reference's operator interface, bound to the C ABI.

  Plan       <- io::load_fluorescences' result-key precomputation (reference src/io/parser.cu:68-154)
  Engine     <- the device-resident state the reference rebuilds on every call
                (src/simulation/proliferation.cu:35-73, cells_population.cu:27-40)
  proliferate()  <- simulation::create_cells_population + simulation::proliferate
                (src/simulation/cells_population.h:12-18, proliferation.h:12-19)
  Simulator  <- simulation::Simulator's four-step lifecycle (src/simulation/simulator.h:10-36):
                load_params / create_cell_population / start_simulation / save_results

Everything that simulates runs the sm_100a kernels through libprocell_b200.so; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (KERNEL_COOP, KERNEL_SIMPLE, SEEDING_IDEAL, SEEDING_REFCOMPAT, CellType, ProcellError,  # noqa: F401
                   RunStats, SimParams, check)

_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def _types_array(types) -> np.ndarray:
    t = np.ascontiguousarray(types, dtype=np.float64)
    if t.ndim == 2:
        t = t[None]
    if t.ndim != 3 or t.shape[2] != 3:
        raise ValueError("types must be [n_types][3] or [n_sets][n_types][3] (proportion, mean, stddev)")
    return t


def read_histogram(path: str):
    lib = _lib.load()
    v, f, n = _f64p(), _u64p(), C.c_size_t()
    check(lib.procell_read_histogram(str(path).encode(), C.byref(v), C.byref(f), C.byref(n)))
    try:
        values = np.ctypeslib.as_array(v, shape=(max(n.value, 1),))[: n.value].copy()
        freqs = np.ctypeslib.as_array(f, shape=(max(n.value, 1),))[: n.value].copy()
    finally:
        lib.procell_free(v)
        lib.procell_free(f)
    return values, freqs


def parse_histogram(text: bytes):
    """The histogram reader on text already in memory."""
    lib = _lib.load()
    text = bytes(text)
    v, f, n = _f64p(), _u64p(), C.c_size_t()
    check(lib.procell_parse_histogram(text, len(text), C.byref(v), C.byref(f), C.byref(n)))
    try:
        values = np.ctypeslib.as_array(v, shape=(max(n.value, 1),))[: n.value].copy()
        freqs = np.ctypeslib.as_array(f, shape=(max(n.value, 1),))[: n.value].copy()
    finally:
        lib.procell_free(v)
        lib.procell_free(f)
    return values, freqs


def parse_cell_types(text: bytes) -> np.ndarray:
    """The cell-types reader on text already in memory (checks the proportion sum like the file reader)."""
    lib = _lib.load()
    text = bytes(text)
    t, n = C.POINTER(CellType)(), C.c_size_t()
    rc = lib.procell_parse_cell_types(text, len(text), C.byref(t), C.byref(n))
    try:
        check(rc)
        return np.array([(t[i].proportion, t[i].mean, t[i].stddev) for i in range(n.value)], dtype=np.float64).reshape(-1, 3)
    finally:
        lib.procell_free(t)


def read_cell_types(path: str) -> np.ndarray:
    lib = _lib.load()
    t, n = C.POINTER(CellType)(), C.c_size_t()
    rc = lib.procell_read_cell_types(str(path).encode(), C.byref(t), C.byref(n))
    try:
        check(rc)
        return np.array([(t[i].proportion, t[i].mean, t[i].stddev) for i in range(n.value)], dtype=np.float64)
    finally:
        lib.procell_free(t)


def write_histogram(path: Optional[str], row_value, row_freq, row_ratio=None) -> None:
    lib = _lib.load()
    rv = np.ascontiguousarray(row_value, dtype=np.float64)
    rf = np.ascontiguousarray(row_freq, dtype=np.int64)
    rr = None if row_ratio is None else np.ascontiguousarray(row_ratio, dtype=np.int64)
    n_types = 0 if rr is None else rr.shape[1]
    check(lib.procell_write_histogram(None if path is None else str(path).encode(), int(rr is not None), n_types,
                                      len(rv), rv.ctypes.data_as(_f64p), rf.ctypes.data_as(_i64p),
                                      None if rr is None else rr.ctypes.data_as(_i64p)))


class Plan:
    """Key space (bin, k) and merged output rows of one input histogram."""

    def __init__(self, values, freqs, phi: float = 0.0):
        lib = _lib.load()
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.freqs = np.ascontiguousarray(freqs, dtype=np.uint64)
        h = C.c_void_p()
        check(lib.procell_plan_create(self.values.ctypes.data_as(_f64p), self.freqs.ctypes.data_as(_u64p),
                                      len(self.values), float(phi), C.byref(h)))
        self.h = h
        self.n_bins = lib.procell_plan_n_bins(h)
        self.n_keys = lib.procell_plan_n_keys(h)
        self.n_rows = lib.procell_plan_n_rows(h)
        self.n_cells = lib.procell_plan_n_cells(h)
        self.phi = lib.procell_plan_phi(h)
        self.depth_capped = bool(lib.procell_plan_depth_capped(h))
        self.row_value = np.zeros(self.n_rows, dtype=np.float64)
        self.key_row = np.zeros(self.n_keys, dtype=np.uint32)
        self.bin_keybase = np.zeros(self.n_bins, dtype=np.uint32)
        self.bin_kdiv = np.zeros(self.n_bins, dtype=np.uint8)
        check(lib.procell_plan_export(h, self.row_value.ctypes.data_as(_f64p), self.key_row.ctypes.data_as(_u32p),
                                      self.bin_keybase.ctypes.data_as(_u32p), self.bin_kdiv.ctypes.data_as(_u8p)))

    @classmethod
    def from_file(cls, path: str, phi: float = 0.0) -> "Plan":
        v, f = read_histogram(path)
        return cls(v, f, phi)

    def lineage_depth(self, types, t_max: float) -> float:
        """Expected depth of a lineage tree for one parameter set: min(t_max / fastest mean, mean halvings phi allows)
        (procell_plan_lineage_depth; the library picks the deep-tree kernel instance from it)."""
        t = _types_array(types)[0]
        return float(_lib.load().procell_plan_lineage_depth(self.h, t.ctypes.data_as(_f64p), len(t), float(t_max)))

    def merge_rows(self, counts_one_set: np.ndarray):
        """counts [n_keys][n_types] -> (row_freq [n_rows], row_ratio [n_rows][n_types])."""
        c = np.ascontiguousarray(counts_one_set, dtype=np.int64)
        n_types = c.shape[1]
        rf = np.zeros(self.n_rows, dtype=np.int64)
        rr = np.zeros((self.n_rows, n_types), dtype=np.int64)
        check(_lib.load().procell_merge_rows(self.h, c.ctypes.data_as(_i64p), n_types, rf.ctypes.data_as(_i64p),
                                             rr.ctypes.data_as(_i64p)))
        return rf, rr

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.load().procell_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _make_params(types: np.ndarray, t_max, seed, seeding_mode, kernel, shard, checkpoints=None, shard_level=0):
    n_sets, n_types, _ = types.shape
    p = SimParams()
    if checkpoints is not None:
        cp = np.ascontiguousarray(checkpoints, dtype=np.float64)
        p._checkpoints_keepalive = cp          # the struct only holds a pointer
        p.checkpoints = cp.ctypes.data_as(C.POINTER(C.c_double))
        p.n_checkpoints = len(cp)
    p.types = types.ctypes.data_as(C.POINTER(CellType))
    p.n_types = n_types
    p.n_sets = n_sets
    p.t_max = float(t_max)
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    p.seeding_mode = int(seeding_mode)
    p.kernel = int(kernel)
    p.shard_rank, p.shard_world, p.shard_unit = int(shard[0]), int(shard[1]), int(shard[2])
    p.shard_level = int(shard_level)
    return p


@dataclass
class Result:
    counts: np.ndarray       # [n_sets][n_keys][n_types] int64
    divisions: np.ndarray    # [n_sets] int64
    stats: dict


def _stats_dict(st: RunStats) -> dict:
    return dict(divisions=int(st.divisions), kernel_ms=float(st.kernel_ms), n_launches=int(st.n_launches),
                grid=int(st.grid), block=int(st.block), smem_bytes=int(st.smem_bytes), donations=int(st.donations),
                seed_phase_us=float(st.seed_phase_us), total_us=float(st.total_us),
                idle_warp_us=float(st.idle_warp_us), idle_waits=int(st.idle_waits))


def proliferate(plan: Plan, types, t_max: float, seed: int = 0x5EED0000, seeding_mode: int = SEEDING_IDEAL,
                kernel: int = KERNEL_COOP, device: int = 0, shard=(0, 1, 0), checkpoints=None, shard_level: int = 0) -> Result:
    """One-shot host-buffer call (what the CLI uses).  shard = (rank, world, unit).  With `checkpoints` (ascending,
    at most 8; the last replaces t_max) counts gets a leading checkpoint axis.  shard_level >= 1 (with world > 1)
    shards subtrees at that tree level instead of whole lineages (procell_sim_params.shard_level)."""
    lib = _lib.load()
    t = _types_array(types)
    p = _make_params(t, t_max, seed, seeding_mode, kernel, shard, checkpoints, shard_level)
    shape = (t.shape[0], plan.n_keys, t.shape[1])
    if checkpoints is not None:
        shape = (len(checkpoints),) + shape
    flat = np.zeros(max(int(np.prod(shape)), 1), dtype=np.int64)     # never a NULL pointer, even for 0 keys
    div = np.zeros(t.shape[0], dtype=np.int64)
    st = RunStats()
    check(lib.procell_proliferate(plan.h, C.byref(p), int(device), flat.ctypes.data_as(_i64p),
                                  div.ctypes.data_as(_i64p), C.byref(st)))
    return Result(flat[: int(np.prod(shape))].reshape(shape), div, _stats_dict(st))


def proliferate_multi(plan: Plan, types, t_max: float, seed: int = 0x5EED0000, n_gpus: int = 0,
                      seeding_mode: int = SEEDING_IDEAL, kernel: int = KERNEL_COOP, shard_level: int = 0) -> Result:
    """One process, n_gpus GPUs of this box (0 = all): sharded seed-cell units (or, with shard_level >= 1, subtrees at
    that tree level) + one NCCL reduce onto GPU 0."""
    lib = _lib.load()
    t = _types_array(types)
    p = _make_params(t, t_max, seed, seeding_mode, kernel, (0, 1, 0), None, shard_level)
    shape = (t.shape[0], plan.n_keys, t.shape[1])
    flat = np.zeros(max(int(np.prod(shape)), 1), dtype=np.int64)
    div = np.zeros(t.shape[0], dtype=np.int64)
    st = RunStats()
    check(lib.procell_proliferate_multi(plan.h, C.byref(p), int(n_gpus), flat.ctypes.data_as(_i64p),
                                        div.ctypes.data_as(_i64p), C.byref(st)))
    return Result(flat[: int(np.prod(shape))].reshape(shape), div, _stats_dict(st))


class Engine:
    """Device-resident engine: tables stay in HBM between runs; run() is asynchronous on a CUDA stream."""

    def __init__(self, device: int = 0):
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.procell_engine_create(int(device), C.byref(h)))
        self.h = h
        self.device = device
        self.plan = None
        self.shape = None

    def load(self, plan: Plan, types, t_max: float, seed: int = 0x5EED0000, seeding_mode: int = SEEDING_IDEAL,
             kernel: int = KERNEL_COOP, shard=(0, 1, 0), checkpoints=None, shard_level: int = 0) -> None:
        t = _types_array(types)
        p = _make_params(t, t_max, seed, seeding_mode, kernel, shard, checkpoints, shard_level)
        check(_lib.load().procell_engine_load(self.h, plan.h, C.byref(p)))
        self.plan = plan
        self.seed = seed
        self.shape = (t.shape[0], plan.n_keys, t.shape[1])
        if checkpoints is not None:
            self.shape = (len(checkpoints),) + self.shape

    def run(self, seed: Optional[int] = None, stream: int = 0, d_counts: int = 0, d_divisions: int = 0) -> None:
        """stream: cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream); d_counts / d_divisions:
        device pointers of int64 tensors shaped [n_sets][n_keys][n_types] / [n_sets], 0 = engine-owned."""
        s = self.seed if seed is None else seed
        check(_lib.load().procell_engine_run(self.h, int(s) & 0xFFFFFFFFFFFFFFFF, C.c_void_p(stream or None),
                                             C.c_void_p(d_counts or None), C.c_void_p(d_divisions or None)))

    def finish(self, stream: int = 0, fetch: bool = True) -> Result:
        st = RunStats()
        n = int(np.prod(self.shape))
        flat = np.zeros(max(n, 1), dtype=np.int64) if fetch else None
        div = np.zeros(self.shape[-3], dtype=np.int64)
        check(_lib.load().procell_engine_finish(self.h, C.c_void_p(stream or None),
                                                flat.ctypes.data_as(_i64p) if fetch else None,
                                                div.ctypes.data_as(_i64p), C.byref(st)))
        return Result(flat[:n].reshape(self.shape) if fetch else None, div, _stats_dict(st))

    def set_target(self, values, freqs) -> None:
        """Target histogram of a calibration sweep (channel values strictly ascending)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        f = np.ascontiguousarray(freqs, dtype=np.uint64)
        check(_lib.load().procell_engine_set_target(self.h, v.ctypes.data_as(_f64p), f.ctypes.data_as(_u64p), len(v)))

    def fitness(self, stream: int = 0, d_counts: int = 0) -> np.ndarray:
        """Hellinger distance of every parameter set's simulated histogram to the target, computed on the GPU."""
        out = np.zeros(self.shape[-3], dtype=np.float64)
        check(_lib.load().procell_engine_fitness(self.h, C.c_void_p(stream or None), C.c_void_p(d_counts or None),
                                                 out.ctypes.data_as(_f64p)))
        return out

    def fitness_in_launch(self) -> bool:
        """True if the last fitness() was computed by the simulation launch itself (target set before run())."""
        return bool(_lib.load().procell_engine_fitness_in_launch(self.h))

    def kernel_mode(self) -> int:
        """Instance of the simulation kernel the loaded simulation runs on (0 base, 1 subtree sharding, 2 set-relative
        sweep table, 3 deep trees with merged leaf counts, -1 bring-up kernel): procell_engine_kernel_mode."""
        return int(_lib.load().procell_engine_kernel_mode(self.h))

    def close(self):
        if getattr(self, "h", None):
            _lib.load().procell_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def simulate(values, freqs, types, t_max, phi=0.0, seed=0x5EED0000, track_ratio=True, n_gpus=1,
             seeding_mode=SEEDING_IDEAL):
    """procell_simulate: histogram arrays in, result rows out, one C call (plan + simulation + merge).
    Returns (row_value[n_rows], freq[n_sets][n_rows], ratio[n_sets][n_rows][n_types] or None, divisions, kernel_ms)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    f = np.ascontiguousarray(freqs, dtype=np.uint64)
    t = np.ascontiguousarray(types, dtype=np.float64)
    if t.ndim == 2:
        t = t[None]
    n_sets, n_types = t.shape[0], t.shape[1]
    inp = _lib.SimInput(v.ctypes.data_as(_f64p), f.ctypes.data_as(_u64p), len(v),
                        t.ctypes.data_as(C.POINTER(_lib.CellType)), n_types, n_sets, float(t_max), float(phi),
                        int(bool(track_ratio)), int(seed), int(n_gpus), int(seeding_mode))
    out = _lib.SimOutput()
    lib = _lib.load()
    check(lib.procell_simulate(C.byref(inp), C.byref(out)))
    try:
        n = out.n_rows
        rows = np.ctypeslib.as_array(out.value, shape=(n,)).copy() if n else np.zeros(0)
        freq = np.ctypeslib.as_array(out.freq, shape=(n_sets, n)).copy() if n else np.zeros((n_sets, 0), np.int64)
        ratio = None
        if track_ratio:
            ratio = (np.ctypeslib.as_array(out.ratio, shape=(n_sets, n, n_types)).copy() if n
                     else np.zeros((n_sets, 0, n_types), np.int64))
        return rows, freq, ratio, int(out.divisions), float(out.kernel_ms)
    finally:
        lib.procell_output_free(C.byref(out))


def rng_ceiling(device: int = 0, iters: int = 2048):
    """(ms, pairs): time of the RNG-only micro-kernel and the number of divisions' worth of draws (one Philox block + two fast ziggurat tests each) it made."""
    ms, pairs = C.c_double(), C.c_double()
    check(_lib.load().procell_rng_ceiling(int(device), int(iters), C.byref(ms), C.byref(pairs)))
    return ms.value, pairs.value


def rng_ceiling_variants(device: int = 0, iters: int = 2048):
    """[(ms, pairs)] of the three shapes of the RNG-only loop (rng_ceiling reports the fastest)."""
    ms, pairs = (C.c_double * 3)(), (C.c_double * 3)()
    check(_lib.load().procell_rng_ceiling_variants(int(device), int(iters), ms, pairs))
    return [(ms[i], pairs[i]) for i in range(3)]


class CmdArgs:
    """io::CmdArgs (reference src/io/cmdargs.h:13-48): plain fields, no parsing here (the CLI parses)."""

    def __init__(self, h0, cell_types, t_max, phi_min=0.0, output_histogram=None, track_ratio=False,
                 tree_depth=23, seed=0x5EED0000):
        self.h0 = h0
        self.cell_types = cell_types
        self.t_max = float(t_max)
        self.phi_min = float(phi_min)
        self.output_histogram = output_histogram
        self.output_histogram_given = output_histogram is not None
        self.track_ratio = bool(track_ratio)
        self.tree_depth = tree_depth
        self.seed = seed


class Simulator:
    """simulation::Simulator (reference src/simulation/simulator.h:10-36), same four calls in the same order."""

    def __init__(self, args: CmdArgs, device: int = 0, kernel: int = KERNEL_COOP, seeding_mode: int = SEEDING_IDEAL):
        self.args = args
        self.device = device
        self.kernel = kernel
        self.seeding_mode = seeding_mode
        self.plan = None
        self.params = None
        self.result = None
        self.initial_population_size = 0

    def load_params(self) -> None:                      # simulator.cu:10-20
        self.plan = Plan.from_file(self.args.h0, self.args.phi_min)
        self.initial_population_size = self.plan.n_cells
        self.params = read_cell_types(self.args.cell_types)

    def create_cell_population(self) -> None:           # simulator.cu:22-27
        """Seed cells are built inside the simulation kernel (never materialised), so nothing happens here."""

    def start_simulation(self) -> bool:                 # simulator.cu:29-38
        self.result = proliferate(self.plan, self.params, self.args.t_max, self.args.seed, self.seeding_mode,
                                  self.kernel, self.device)
        return True

    def save_results(self) -> None:                     # simulator.cu:40-56
        rf, rr = self.plan.merge_rows(self.result.counts[0])
        write_histogram(self.args.output_histogram if self.args.output_histogram_given else None,
                        self.plan.row_value, rf, rr if self.args.track_ratio else None)
