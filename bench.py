#!/usr/bin/env python3
"""bench.py - simulated cell divisions per second of the proliferation hot path on N B200s of one node.

Workload of the headline (`value`, `e2e`): BASELINE.json configs[1] - 1e6 seed cells per GPU, 3 proliferating types +
quiescent, t_max = 240, -r per-type counts, phi = 0.5; ~1.1e8 divisions per simulation and GPU.
A STEP is one batch of B such simulations (--batch, default 32; each with its own Philox seed, each its own launch
and - at N > 1 - its own NCCL reduce), so that the driver's 20 steps cover ~0.7 s of GPU time instead of 22 ms.
N > 1 is weak scaling: the histogram holds N x 1e6 cells, seed-cell units are sharded rank-strided, every simulation
ends with ONE NCCL reduce (sum, int64) of count tensor + division counter to rank 0; the reduce of simulation i runs on
a second stream and overlaps simulation i + 1.  Every simulation of the timed region keeps its own result tensor (a pool
in HBM), so nothing is accumulated or overwritten inside the timed region.

  value       whole-job divisions/s, tables resident in HBM: CUDA-event time of the K steps (L2 flushed between steps,
              outside the events), max over ranks
  e2e         the same metric through the C ABI with HOST buffers: histogram arrays -> plan -> H2D tables -> kernel ->
              [reduce] -> D2H count tensor -> merged rows, wall clock bracketed by synchronize, two engines in flight
  per_config  BASELINE configs 1..5 at full size (and the 1e5-cell config-2 shape the reference arm can follow), each with
              its own ms, divisions/s and roofline fraction - by divisions and by draws; at N > 1 each input is sharded over
              the ranks, by seed-cell units or - config 4 - by subtrees (strong scaling), and checked against rank 0's own run
  roofline    instruction-issue / shared-memory roofline (neither HBM nor tensor bound - DESIGN.md section 5):
              achieved divisions/s over the RNG-only ceiling kernel measured in the same run, with the hardware-unit
              fractions of the committed ncu capture beside it; HBM figures for completeness
  cpu_baseline  the CPU oracle (oracle/, a port with the same Philox streams) on this box's host cores
  verify      N > 1: the reduced tensor of one simulation equals, bit for bit, the same simulation run unsharded on rank 0

--impl reference times the reference's own CUDA build (oracle/_ref/, one B200) on the SAME 1e6-cell input, run as ten
processes of 1e5 cells because one reference process cannot simulate more on sm_100; see run_reference().
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "simulated cell divisions/sec"
UNIT = "divisions/s"
CELLS_PER_GPU = 1_000_000
FLUSH_BYTES = 256 << 20
SHARD_UNIT = 32
WORKLOAD_TEXT = ("BASELINE configs[1]: 1e6 seed cells per GPU (synthetic 1024-channel histogram), types 0.40/48.33/21.6 "
                 "0.25/86.3/26.8 0.17/24/6 + 0.18 quiescent, t_max=240, phi=0.5, -r")


def workload_for(n_gpus: int):
    from cuda_pro_cell_b200 import synth
    w = synth.workload(2)
    if n_gpus > 1:
        w.values, w.freqs = synth.synthetic_histogram(CELLS_PER_GPU * n_gpus)
        w.n_cells = CELLS_PER_GPU * n_gpus
    return w


class ClockSampler:
    """SM clock and clock-event (throttle) reasons of one GPU sampled every 20 ms through NVML from a thread, started
    BEFORE warm-up; stop() summarises the samples that fell inside the timed region [t0, t1].  Falls back to an
    `nvidia-smi -lms 100` child process when pynvml is not importable."""
    NAMES = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
             "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}

    def __init__(self, index: int):
        self.index = index
        self.samples = []          # (perf_counter, sm_mhz, reason bits)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.smi = None
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {k: getattr(pynvml, v) for k, v in self.NAMES.items()}

            def pump():
                while not self.stop_flag.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        rs = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.samples.append((time.perf_counter(), mhz, [k for k, b in bits.items() if rs & b]))
                    except Exception:
                        pass
                    self.stop_flag.wait(0.02)

            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
            self.how = "pynvml, 20 ms"
        except Exception:
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.smi = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                             "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

                def pump_smi():
                    for ln in self.smi.stdout:
                        f = [x.strip() for x in ln.split(",")]
                        try:
                            self.max_mhz = float(f[1])
                            self.samples.append((time.perf_counter(), float(f[0]),
                                                 [k for k, v in zip(self.NAMES, f[2:6]) if v.lower().startswith("active")]))
                        except (ValueError, IndexError):
                            pass

                self.thread = threading.Thread(target=pump_smi, daemon=True)
                self.thread.start()
                self.how = "nvidia-smi -lms 100"
            except OSError:
                self.how = None

    def stop(self, t0: float, t1: float):
        self.stop_flag.set()
        if self.smi is not None:
            self.smi.terminate()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        reasons = sorted({r for s in inside for r in s[2]})
        return {"sm_mhz": statistics.median(s[1] for s in inside) if inside else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(inside), "samples_total": len(self.samples), "how": self.how}


def cpu_baseline(w, max_seconds=25.0):
    """The oracle port on this box's host cores, on the same workload (full config-2 shard of one GPU, repeated
    with fresh seeds until ~10 s of CPU work)."""
    import oracle_lib
    from cuda_pro_cell_b200 import synth
    values, freqs = synth.synthetic_histogram(CELLS_PER_GPU)
    plan = oracle_lib.OraclePlan(values, freqs, w.phi)
    cores = oracle_lib.n_host_threads()
    t_total, div_total, runs = 0.0, 0, 0
    while t_total < 10.0 and runs < 200:       # about 10 s of CPU work on all host cores
        t0 = time.perf_counter()
        r = oracle_lib.simulate(plan, w.types, w.t_max, w.seed + runs, n_threads=cores)
        t_total += time.perf_counter() - t0
        div_total += int(r["divisions"].sum())
        runs += 1
        if t_total > max_seconds:
            break
    return {"value": div_total / t_total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d full run(s) of the 1e6-cell config-2 workload (%.3g divisions, %.2f s)" % (runs, div_total, t_total)}


# --------------------------------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------------------------------
def _ref_run(binary, tmp, wl, t_max, limit):
    """one process of a reference build on the files in tmp; returns (wall_s, leaf total) or None on failure"""
    o = tmp / "o.txt"
    if o.exists():
        o.unlink()
    cmd = [str(binary), "-h", str(tmp / "h.txt"), "-c", str(tmp / "c.txt"), "-t", repr(float(t_max)), "-p", repr(float(wl.phi)),
           "-o", str(o)] + (["-r"] if wl.track_ratio else [])
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=limit)
    except subprocess.TimeoutExpired:
        return None
    dt = time.perf_counter() - t0
    if r.returncode != 0 or not o.exists():
        return None
    leaves = sum(int(ln.split("\t")[1]) for ln in o.read_text().splitlines() if ln.strip())
    return dt, leaves


def run_reference(args):
    """--impl reference: the reference's own CUDA build on one B200 on the SAME input as our arm (configs[1], 1e6 cells,
    written to the text files both executables read).

    Which binary: the unmodified build (oracle/_ref/procell_ref) silently loses subtrees on sm_100 from 2e4 cells on
    (CDP2's pending-launch pool; tests/golden/ref_cfg2_*.json), so the arm uses oracle/_ref/procell_ref_pl =
    the same sources plus the ONE line BASELINE.md section 2 permits, cudaDeviceSetLimit(cudaLimitDevRuntimePendingLaunchCount,
    N) (oracle/Makefile `refpl`, diff in oracle/_ref/procell_ref_pl.diff).  With it the reference simulates up to ~1e5 cells
    of this shape correctly; at 1e6 cells in one process it still loses 77 % of the leaves, because its own 16-bit grid
    size (proliferation.cu:245) truncates the dense level arrays.  The arm therefore runs the same 1e6-cell
    input as TEN processes of 1e5 cells (every bin's frequency dealt out over ten histograms of the same shape; cells are
    independent, so the summed output has the law of the single run) - what a user of the reference would have to do.  A
    run is accepted only if its summed leaf total is within 2 % of the Philox oracle's expectation for the whole input.
    What is reported: the reference has no timers and no resident mode, so every simulation is a whole process.  `value` =
    divisions / sum(wall_i - floor), where floor = the same command with -t 0 (context creation, file parsing, seed
    population; no division) - the simulation alone, the quantity our event-timed arm measures; the whole-process figure
    is given beside it (`process_value`), and so is the configs[0] command line, wall clock, for the CLI-vs-CLI
    comparison (`cli_config1`; our arm: per_config.config1.cli_wall_ms).  If even the chunked run fails, the line says so
    (`unavailable_config2`) and carries configs[0] (1e4 cells), which the unmodified reference does run correctly."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle_lib
    from cuda_pro_cell_b200 import synth
    refdir = ROOT / "oracle" / "_ref"
    w = workload_for(1)
    cfg = {"workload": WORKLOAD_TEXT, "n_cells": int(w.n_cells), "l2": "n/a (separate process per step)"}
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "requested_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg}
    binaries = [b for b in (refdir / "procell_ref_pl", refdir / "procell_ref") if b.exists()]
    if not binaries:
        # no reference binary on this box: time the oracle port on all host cores instead
        cb = cpu_baseline(w)
        line.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "note": "oracle/_ref/procell_ref* missing; oracle port timed instead"})
        print(json.dumps(line))
        return 0
    tmp = Path(tempfile.mkdtemp(prefix="procell_ref_"))
    # the arm is the same one-GPU measurement whatever --gpus says (the reference has no multi-GPU path: device 0 is
    # hard-coded, proliferation.cu:38), so the repeats of a scaling run are kept shorter
    budget_s = 170.0 if args.gpus <= 1 else 110.0
    t_begin = time.perf_counter()

    def measure(binary, wl, label, max_steps):
        """floor + up to max_steps timed processes; None unless every run's leaf total is credible"""
        (tmp / "h.txt").write_text(synth.histogram_text(wl.values, wl.freqs))
        (tmp / "c.txt").write_text(synth.types_text(wl.types[0]))
        oplan = oracle_lib.OraclePlan(wl.values, wl.freqs, wl.phi)
        expect = oracle_lib.simulate(oplan, wl.types, wl.t_max, wl.seed)
        exp_div, exp_leaves = int(expect["divisions"].sum()), int(expect["row_freq"].sum())
        floors = []
        for _ in range(2):
            f = _ref_run(binary, tmp, wl, 0.0, 120)
            if f is None:
                return None
            floors.append(f[0])
            time.sleep(max(0.0, 1.05 - f[0]))
        floor = min(floors)
        first = _ref_run(binary, tmp, wl, wl.t_max, 160)         # warm-up step; also sizes the loop
        if first is None or abs(first[1] - exp_leaves) > 0.02 * exp_leaves:
            return {"ok": False, "leaves": None if first is None else first[1], "expected_leaves": exp_leaves}
        left = budget_s - (time.perf_counter() - t_begin)
        n = max(1, min(max_steps, int(left / max(first[0] + 0.2, 1.1))))
        times, leaves = [], []
        for _ in range(n):
            time.sleep(max(0.0, 1.05 - (times[-1] if times else first[0])))   # the reference seeds from time(NULL)
            r = _ref_run(binary, tmp, wl, wl.t_max, 160)
            if r is None or abs(r[1] - exp_leaves) > 0.02 * exp_leaves:
                return {"ok": False, "leaves": None if r is None else r[1], "expected_leaves": exp_leaves}
            times.append(r[0])
            leaves.append(r[1])
        wall = sum(times) / len(times)
        sim = max(wall - floor, 1e-6)
        return {"ok": True, "label": label, "binary": binary.name, "steps": len(times), "wall_s": wall, "floor_s": floor, "sim_s": sim,
                "divisions": exp_div, "leaves": leaves[-3:], "expected_leaves": exp_leaves}

    def measure_chunked(binary, wl, n_chunks, max_steps):
        """The SAME input run as n_chunks processes: every bin's frequency is dealt out over n_chunks histograms of the same
        shape (cells are independent, so the summed output has the law of the single run), because one reference process cannot
        hold the whole input.  A step = the n_chunks processes one after the other; simulation time = sum(wall_i - floor)."""
        oplan = oracle_lib.OraclePlan(wl.values, wl.freqs, wl.phi)
        expect = oracle_lib.simulate(oplan, wl.types, wl.t_max, wl.seed)
        exp_div, exp_leaves = int(expect["divisions"].sum()), int(expect["row_freq"].sum())
        fr = wl.freqs.astype(np.int64)
        chunks = [fr // n_chunks + ((fr % n_chunks) > c).astype(np.int64) for c in range(n_chunks)]
        assert sum(int(c.sum()) for c in chunks) == int(fr.sum())
        (tmp / "c.txt").write_text(synth.types_text(wl.types[0]))

        def one_step(t_max):
            walls, leaves = [], 0
            for c in range(n_chunks):
                (tmp / "h.txt").write_text(synth.histogram_text(wl.values, chunks[c].astype(np.uint64)))
                r = _ref_run(binary, tmp, wl, t_max, 120)
                if r is None:
                    return None
                walls.append(r[0])
                leaves += r[1]
                time.sleep(max(0.0, 1.05 - r[0]))           # the reference seeds from time(NULL)
            return walls, leaves

        (tmp / "h.txt").write_text(synth.histogram_text(wl.values, chunks[0].astype(np.uint64)))
        floors = []
        for _ in range(2):
            f = _ref_run(binary, tmp, wl, 0.0, 120)
            if f is None:
                return None
            floors.append(f[0])
            time.sleep(max(0.0, 1.05 - f[0]))
        floor = min(floors)
        t0 = time.perf_counter()
        first = one_step(wl.t_max)                          # warm-up step; also sizes the loop
        step_s = time.perf_counter() - t0
        if first is None or abs(first[1] - exp_leaves) > 0.02 * exp_leaves:
            return {"ok": False, "leaves": None if first is None else first[1], "expected_leaves": exp_leaves}
        left = budget_s - (time.perf_counter() - t_begin)
        n = max(1, min(max_steps, int(left / (step_s + 0.5))))
        sims, walls_all, leaves_all = [], [], []
        for _ in range(n):
            r = one_step(wl.t_max)
            if r is None or abs(r[1] - exp_leaves) > 0.02 * exp_leaves:
                return {"ok": False, "leaves": None if r is None else r[1], "expected_leaves": exp_leaves}
            sims.append(sum(max(x - floor, 0.0) for x in r[0]))
            walls_all.append(sum(r[0]))
            leaves_all.append(r[1])
        return {"ok": True, "label": "configs[1] full (1e6 cells) as %d processes of 1e5 cells" % n_chunks, "binary": binary.name,
                "steps": len(sims), "wall_s": sum(walls_all) / len(walls_all), "floor_s": floor * n_chunks,
                "sim_s": sum(sims) / len(sims), "divisions": exp_div, "leaves": leaves_all[-3:], "expected_leaves": exp_leaves,
                "chunks": n_chunks}

    # the config-1 command line, wall clock (what the unmodified reference certainly runs): CLI-vs-CLI figure
    w1 = synth.workload(1)
    cli1 = measure(refdir / "procell_ref" if (refdir / "procell_ref").exists() else binaries[0], w1, "configs[0]", 3)
    # configs[1] in ONE process: fails on sm_100 for every build (the unmodified one loses 93 % of the leaves at 1e6 cells,
    # the one with the raised pending-launch pool 77 %: profiles/r2a_ref_probe.json) - the reference's own 16-bit grid size
    # (proliferation.cu:245) caps a launch at 65535 x 1024 cells and its dense level arrays pass that after seven
    # levels.  Attempted once here with the patched build, so that the line carries the evidence.
    result, rejected = None, []
    pl = refdir / "procell_ref_pl"
    if pl.exists():
        m = measure(pl, w, "configs[1] full (1e6 cells), one process", args.steps)
        if m and m.get("ok"):
            result = m
        else:
            rejected.append({"binary": pl.name, "one_process": m})
    # ... so the same input is run the way a user of the reference would have to: as ten processes of 1e5 cells, the largest
    # input of this shape a reference build simulates correctly (leaf total 99.8 % of the oracle's expectation, fluorescence
    # mass conserved; needs the raised pending-launch pool)
    if result is None and pl.exists() and time.perf_counter() - t_begin < budget_s - 60:
        m = measure_chunked(pl, w, 10, args.steps)
        if m and m.get("ok"):
            result = m
            line["one_process_fails"] = rejected
            cfg["how_run"] = ("the 1e6-cell histogram dealt out over 10 processes of 1e5 cells each (same bins, frequencies split), outputs "
                              "summed: one reference process cannot simulate more than ~1e5 cells of this shape on sm_100")
        else:
            rejected.append({"binary": pl.name, "ten_processes": m})
    if result is None and cli1 and cli1.get("ok"):
        result = cli1
        cfg["workload"] = "BASELINE configs[0]: 1e4 seed cells, types 0.53/48.33/21.6 0.29/86.3/26.8 + 0.18 quiescent, t_max=168, phi=min bin"
        cfg["n_cells"] = int(w1.n_cells)
        line["unavailable_config2"] = ("no reference build simulates the 1e6-cell input credibly on sm_100 (16-bit grid size, "
                                       "proliferation.cu:245; pending-launch pool): %s" % json.dumps(rejected))
    if result is None:
        line.update({"unavailable": "reference binary failed on every attempted input: %s" % json.dumps(rejected)})
        print(json.dumps(line))
        return 0
    value = result["divisions"] / result["sim_s"]
    cfg["workload_run"] = result["label"]
    patch = ""
    if result["binary"] != "procell_ref":
        d = refdir / (result["binary"] + ".diff")
        patch = d.read_text().strip() if d.exists() else "cudaDeviceSetLimit(cudaLimitDevRuntimePendingLaunchCount, N)"
    line.update({"value": value, "ms_per_step": 1e3 * result["sim_s"], "steps": result["steps"],
                 "process_value": result["divisions"] / result["wall_s"], "process_wall_ms_per_step": 1e3 * result["wall_s"],
                 "process_start_floor_ms": 1e3 * result["floor_s"],
                 "reference_binary": result["binary"], "reference_patch": patch or "none (unmodified)",
                 "timing": "value = divisions / (process wall - floor); floor = same command with -t 0 (context, parsing, seed population)",
                 "cli_config1": None if not (cli1 and cli1.get("ok")) else
                 {"wall_ms": 1e3 * cli1["wall_s"], "floor_ms": 1e3 * cli1["floor_s"], "divisions": cli1["divisions"], "binary": cli1["binary"]},
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                                  "sample": "%s; reference CUDA build %s (sm_100, CDP2) on one B200, %d process(es); divisions = oracle "
                                            "expectation %d (reference leaves %s vs expected %d)"
                                            % (result["label"], result["binary"], result["steps"], result["divisions"],
                                               result["leaves"], result["expected_leaves"])},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def hw_fractions():
    """hardware-unit fractions of the dominant kernel from the committed ncu captures (profiles/*_hw_fractions.json,
    written by tools/summarize_profile.py --fractions from `ncu --set full` reports of this build)"""
    best = None
    for p in sorted((ROOT / "profiles").glob("*_hw_fractions.json")):
        try:
            best = (p.name, json.loads(p.read_text()))
        except Exception:
            pass
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="simulations per step (each its own launch, seed and reduce)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=8, help="steps (batches of --batch simulations) of the end-to-end leg")
    ap.add_argument("--e2e-engines", type=int, default=2,
                    help="engines (each with its own stream, count tensor and pinned host buffer) the end-to-end leg keeps in "
                         "flight: with 2 the D2H of simulation i and the H2D of simulation i+2 overlap kernel i+1")
    ap.add_argument("--serial-reduce", action="store_true", help="N > 1: reduce on the compute stream (no overlap with the next simulation)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cuda_pro_cell_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    B = max(1, args.batch)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # before warm-up: the first samples exist when the timed region starts

    w = workload_for(world)
    n_types = w.types.shape[1]
    plan = api.Plan(w.values, w.freqs, w.phi)
    shard = (rank, world, SHARD_UNIT)
    # one resident engine on one stream; every simulation of the run writes its OWN result tensor (count tensor + division
    # counter, 141 KB) in a pool in HBM, so nothing is accumulated or overwritten inside the timed region and the division
    # total of the timed steps is read from the pool afterwards.  (Alternating two engines on two streams was measured
    # and changes nothing - 97.58 vs 97.59 G div/s: a persistent kernel's CTAs all retire together - so it is gone.)
    eng = api.Engine(local_rank)
    eng.load(plan, w.types, w.t_max, w.seed, shard=shard)
    n_counts = plan.n_keys * n_types
    n_pool = K * B
    if n_pool * (n_counts + 1) * 8 > (8 << 30):
        raise SystemExit("result pool of %d simulations does not fit: lower --steps or --batch" % n_pool)
    pool = torch.zeros((n_pool, n_counts + 1), dtype=torch.int64, device=dev)
    warm = torch.zeros((B, n_counts + 1), dtype=torch.int64, device=dev)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    comm = torch.cuda.Stream(device=dev) if world > 1 and not args.serial_reduce else None

    def simulate(seed, buf):
        """one simulation into buf; N > 1: its reduce runs on the communication stream and overlaps the next simulation"""
        eng.run(seed, stream.cuda_stream, buf.data_ptr(), buf.data_ptr() + 8 * n_counts)
        if world > 1 and comm is not None:
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
        elif world > 1:
            dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)

    def step(first_seed, bufs):
        if comm is not None:
            comm.wait_stream(stream)         # the warm-up pool is reused: its last reduce must not overlap a new run into it
        for j in range(B):
            simulate(first_seed + j, bufs[j])
        if comm is not None:
            stream.wait_stream(comm)         # a step ends when its last reduce has ended

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        flush.fill_(i)
        step(w.seed + 100000 * i, warm)
    barrier()
    eng.finish(stream.cuda_stream, fetch=False)      # status word (sticky on the device): no failure during warm-up

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xFF)                      # L2 flush (256 MiB > 126 MB L2), outside the event pair
        evs[i][0].record(stream)
        step(w.seed + 1000 + i * B, pool[i * B:(i + 1) * B])
        evs[i][1].record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    t_wall = t_wall1 - t_wall0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    gpu_ms = sum(a.elapsed_time(b) for a, b in evs)
    t_ms = torch.tensor([gpu_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    gpu_ms = float(t_ms.item())
    # the device status word is sticky: a pool overflow / watchdog abort in ANY simulation since the last finish is still there
    eng.finish(stream.cuda_stream, fetch=False)
    div_timed = int(pool[:, n_counts].sum().item())       # rank 0 holds the reduced totals
    div_per_sim = div_timed / (K * B)
    value = div_timed / (gpu_ms * 1e-3) if rank == 0 else 0.0

    # ---- N > 1: the reduced tensor of one simulation == the same simulation unsharded on rank 0 (outside any timed region)
    verify = None
    if world > 1:
        simulate(w.seed + 77, warm[0])
        barrier()
        if rank == 0:
            whole = api.Engine(local_rank)
            whole.load(plan, w.types, w.t_max, w.seed, shard=(0, 1, SHARD_UNIT))
            ref = torch.zeros_like(warm[0])
            whole.run(w.seed + 77, stream.cuda_stream, ref.data_ptr(), ref.data_ptr() + 8 * n_counts)
            whole.finish(stream.cuda_stream, fetch=False)
            verify = {"reduced_tensor_equals_single_gpu_run": bool(torch.equal(ref, warm[0])), "divisions": int(ref[n_counts].item()),
                      "what": "count tensor + division counter of one %d-rank simulation after the NCCL reduce vs the same seed "
                              "run unsharded on rank 0" % world}
            whole.close()
        barrier()

    # ---- end to end through the C ABI with host buffers
    E = max(1, args.e2e_steps) * B           # simulations of the end-to-end leg
    h2d = plan.n_bins * 4 * 2 + plan.n_bins + 4 + w.types.size // 3 * (8 + 1 + 16) + 256 * 8
    d2h = (n_counts + 1) * 8
    host_values, host_freqs = w.values.copy(), w.freqs.copy()
    e2e_div = 0
    n_eng = max(1, args.e2e_engines)
    engs = ([eng] + [api.Engine(local_rank) for _ in range(n_eng)])[:n_eng]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_eng)]
    dbufs = [torch.zeros(n_counts + 1, dtype=torch.int64, device=dev) for _ in range(n_eng)]
    hosts = [torch.zeros(n_counts + 1, dtype=torch.int64).pin_memory() for _ in range(n_eng)]
    inflight = [None] * n_eng

    def retire(k):
        """slot k: wait for its D2H, check the device status word, merge the rows (what a caller reads)"""
        nonlocal e2e_div
        p_k, ev_k = inflight[k]
        ev_k.synchronize()
        engs[k].finish(streams[k].cuda_stream, fetch=False)
        if rank == 0:
            p_k.merge_rows(hosts[k][:n_counts].numpy().reshape(plan.n_keys, n_types))
            e2e_div += int(hosts[k][n_counts])
        inflight[k] = None

    barrier()
    t0 = time.perf_counter()
    # every simulation does all of its own work - plan from the host arrays, H2D tables, kernel, reduce, D2H, row merge; a
    # slot (engine, stream, device tensor, pinned host tensor) is reused only after its previous simulation has been read
    # back, checked and merged, so with two slots the host legs and copies of one overlap the kernel of the other
    for i in range(E):
        k = i % n_eng
        if inflight[k] is not None:
            retire(k)
        p = api.Plan(host_values, host_freqs, w.phi)                      # parser.cu:68-154 work, on the host
        engs[k].load(p, w.types, w.t_max, w.seed + i, shard=shard)        # H2D tables (asynchronous, pinned staging)
        engs[k].run(w.seed + 2000 + i, streams[k].cuda_stream, dbufs[k].data_ptr(), dbufs[k].data_ptr() + 8 * n_counts)
        with torch.cuda.stream(streams[k]):
            if world > 1:
                dist.reduce(dbufs[k], dst=0, op=dist.ReduceOp.SUM)
            hosts[k].copy_(dbufs[k], non_blocking=True)                   # D2H count tensor + division counter
            ev = torch.cuda.Event()
            ev.record(streams[k])
        inflight[k] = (p, ev)
    for j in range(n_eng):
        k = (E + j) % n_eng
        if inflight[k] is not None:
            retire(k)
    barrier()
    t_e2e = time.perf_counter() - t0
    for extra in engs:
        extra.close()
    t_e = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e.item())

    # ---- RNG-only ceiling kernel, measured now on this GPU (rank 0)
    ceiling, ceiling_variants = None, None
    if rank == 0:
        # the loop exists in three shapes; the roofline is quoted against the FASTEST one measured now
        names = ["1 chain, 48 warps/SM", "1 chain, 64 warps/SM (32 registers)", "2 chains, 32 warps/SM"]
        var = api.rng_ceiling_variants(local_rank, 4096)
        ceiling_variants = {n: pairs / (ms * 1e-3) / 1e9 for n, (ms, pairs) in zip(names, var)}
        ceiling = max(ceiling_variants.values()) * 1e9

    # ---- per-config records: BASELINE configs 2..5 at full size (strong scaling over the ranks at N > 1)
    per_config = None
    if not args.no_per_config:
        per_config = run_per_config(api, synth, torch, dist, dev, local_rank, rank, world, stream, barrier, ceiling)

    if rank == 0:
        per_gpu = value / world
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = float(h2d + d2h)          # tables read once + count tensor written once per launch
        ms_sim = gpu_ms / (K * B)
        hw = hw_fractions()
        c2 = (hw[1].get("config2") or {}) if hw is not None else {}
        c4 = (hw[1].get("config4") or {}) if hw is not None else {}
        roofline = {"bound": "shared-memory wavefronts + instruction issue",
                    "bound_detail": "shared-memory wavefronts (LSU data pipe: %.2f of its peak on this workload, %.2f on config 4) and "
                                    "instruction issue (%.2f / %.2f of the issue slots) together; neither HBM nor tensor bound "
                                    "(DESIGN.md section 5, roofline.hardware)"
                                    % (c2.get("shared_wavefront_frac", float("nan")), c4.get("shared_wavefront_frac", float("nan")),
                                       c2.get("issue_slot_frac", float("nan")), c4.get("issue_slot_frac", float("nan"))),
                    "binding_unit": {"name": "l1tex data pipe, shared-memory wavefronts", "frac": c2.get("shared_wavefront_frac"),
                                     "issue_slot_frac": c2.get("issue_slot_frac"), "from": "committed ncu capture of this build"},
                    "achieved": per_gpu / 1e9, "peak": ceiling / 1e9, "unit": "Gdivisions/s per GPU",
                    "frac": per_gpu / ceiling,
                    "peak_source": "k_rng_ceiling measured live: the arithmetic of the common DIVIDE iteration alone - one Philox4x32-10 block, "
                                   "two fast ziggurat tests, 2 timers, 2 time updates, the compares - with no tree, ring or atomics; "
                                   "the fastest of three shapes of that loop, with every arithmetic shortcut of the product kernel.  "
                                   "The ceiling rose from 175 (Box-Muller, first half of round 2) to ~270 G/s with the ziggurat draw "
                                   "and the cheaper block / fast-test forms: frac fell although the kernel got faster",
                    "peak_variants_Gdiv_s": ceiling_variants,
                    "traffic": None,
                    "hbm": {"achieved": alg_bytes / (ms_sim * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": alg_bytes / (ms_sim * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                            "algorithmic_bytes_per_launch": alg_bytes}}
        if hw is not None:
            # hardware-unit fractions of k_proliferate_coop on this workload and of the ceiling kernel itself, from the
            # committed `ncu --set full` capture of this build (not measured live: a profiler cannot run inside a bench)
            roofline["hardware"] = dict(hw[1], source="profiles/" + hw[0])
            roofline["traffic"] = hw[1].get("config2", {}).get("dram_bytes_per_launch") if world == 1 else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": gpu_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD_TEXT, "n_cells": int(plan.n_cells), "simulations_per_step": B,
                           "divisions_per_simulation": div_per_sim, "ms_per_simulation": ms_sim,
                           "results": "every simulation of the timed region keeps its own count tensor (pool of %d x %d B in HBM)" % (n_pool, (n_counts + 1) * 8),
                           "sharding": "seed-cell units of %d, rank-strided; one NCCL reduce(sum,int64) per simulation%s"
                                       % (SHARD_UNIT, "" if comm is None else ", on a second stream (overlaps the next simulation)"),
                           "l2": "flushed between steps (256 MiB write), outside the timed events"},
                "wall_ms_per_step_incl_flush": 1e3 * t_wall / K, "timed_region_s": gpu_ms * 1e-3,
                "e2e": {"value": e2e_div / t_e2e if t_e2e > 0 else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * B,
                        "d2h_bytes_per_step": int(d2h) * B, "steps": E // B, "simulations": E,
                        "h2d_bytes_per_simulation": int(h2d), "d2h_bytes_per_simulation": int(d2h),
                        "engines_in_flight": n_eng,
                        "path": "api.Plan (host) -> procell_engine_load (H2D, pinned staging) -> procell_engine_run -> reduce -> D2H -> "
                                "merge_rows, per simulation; %d engine(s) in flight" % n_eng},
                "gpu_launches": 2 * K * B, "kernels_per_simulation": ["k_queue_init", "k_proliferate_coop"],
                "clocks": clocks, "roofline": roofline}
        if verify is not None:
            line["verify"] = verify
        if per_config is not None:
            line["per_config"] = per_config
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is timed on rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cli_wall_config1(synth, w):
    """the `procell` command line on configs[0], whole-process wall clock (what bench.py --impl reference reports for the
    reference binary as cli_config1), and the same command with -t 0 (process start, CUDA context, parsing: no division)"""
    from cuda_pro_cell_b200 import _lib
    tmp = Path(tempfile.mkdtemp(prefix="procell_cli_"))
    (tmp / "h.txt").write_text(synth.histogram_text(w.values, w.freqs))
    (tmp / "c.txt").write_text(synth.types_text(w.types[0]))

    def run(t_max):
        cmd = [str(_lib.CLI_PATH), "-h", str(tmp / "h.txt"), "-c", str(tmp / "c.txt"), "-t", repr(float(t_max)), "-p", repr(float(w.phi)),
               "-o", str(tmp / "o.txt")]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
        return time.perf_counter() - t0 if r.returncode == 0 else None

    floors = [run(0.0) for _ in range(2)]
    walls = [run(w.t_max) for _ in range(3)]
    if any(x is None for x in floors + walls):
        return {"cli_wall_ms": None}
    return {"cli_wall_ms": 1e3 * sum(walls) / len(walls), "cli_floor_ms": 1e3 * min(floors),
            "cli_note": "`procell -h -c -t 168 -p -o`, whole process incl. CUDA context creation, mean of 3; floor = same with -t 0"}


def run_per_config(api, synth, torch, dist, dev, local_rank, rank, world, stream, barrier, ceiling):
    """BASELINE configs 2..5 at full size, one record each: ms per simulation (CUDA events, max over ranks), divisions/s,
    fraction of the live RNG ceiling - by divisions, and by DRAWS (seed cells + divisions: every seed cell also costs a
    Philox block and a timer draw, and configs 3 and 5 are seed-heavy).  N > 1 is STRONG scaling here: the one input
    is sharded over the ranks - seed-cell units (configs 2, 3, 5; unit chosen by the library) or subtrees at tree level 6
    (config 4: a hundred fast lineages own all the work) - and rank 0 checks the reduced tensor against the same simulation
    unsharded on its own GPU."""
    out = {}
    plans = [("config1", synth.workload(1), 3, 50, 0), ("config2", synth.workload(2), 1, 10, 0),
             ("config2_shape_1e5_cells", synth.workload(2, 0.1), 3, 20, 0), ("config3", synth.workload(3), 1, 10, 0),
             ("config4", synth.workload(4), 1, 2, 6 if world > 1 else 0), ("config5", synth.workload(5), 1, 4, 0)]
    for name, w, n_warm, n_timed, level in plans:
        plan = api.Plan(w.values, w.freqs, w.phi)
        n_sets, n_types = w.types.shape[0], w.types.shape[1]
        n_counts = n_sets * plan.n_keys * n_types
        eng = api.Engine(local_rank)
        eng.load(plan, w.types, w.t_max, w.seed, shard=(rank, world, 0), shard_level=level)
        buf = torch.zeros(n_counts + n_sets, dtype=torch.int64, device=dev)

        def sim():
            eng.run(w.seed, stream.cuda_stream, buf.data_ptr(), buf.data_ptr() + 8 * n_counts)
            if world > 1:
                dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)

        for _ in range(n_warm):
            sim()
        barrier()
        evs = []
        for _ in range(n_timed):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            sim()
            b.record(stream)
            evs.append((a, b))
        barrier()
        st = eng.finish(stream.cuda_stream, fetch=False).stats
        ms = sum(a.elapsed_time(b) for a, b in evs) / n_timed
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        idle = torch.tensor([st["idle_warp_us"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(idle, op=dist.ReduceOp.MAX)
        rec = None
        if rank == 0:
            divisions = int(buf[n_counts:].sum().item())
            draws = divisions + int(plan.n_cells) * n_sets
            rec = {"n_cells": int(plan.n_cells), "n_sets": n_sets, "t_max": w.t_max, "divisions": divisions, "ms": ms,
                   "value": divisions / (ms * 1e-3), "unit": UNIT, "scaling": "strong" if world > 1 else "single GPU",
                   "sharding": "single GPU" if world == 1 else ("subtrees at tree level %d" % level if level else "seed-cell units, rank-strided"),
                   "smem_bytes": st["smem_bytes"], "grid": st["grid"], "idle_warp_us_max_rank": float(idle.item()),
                   "rank0_seed_phase_us": st["seed_phase_us"], "rank0_kernel_span_us": st["total_us"], "rank0_donated_chunks": st["donations"],
                   "warps": st["grid"] * st["block"] // 32,
                   "roofline": {"frac": divisions / (ms * 1e-3) / world / ceiling,
                                "frac_draws": draws / (ms * 1e-3) / world / ceiling,
                                "peak": ceiling / 1e9, "unit": "G/s per GPU",
                                "note": "frac = divisions/s per GPU over the live RNG ceiling; frac_draws counts seed cells too"}}
            if world > 1:
                whole = api.Engine(local_rank)
                whole.load(plan, w.types, w.t_max, w.seed)
                ref = torch.zeros_like(buf)
                whole.run(w.seed, stream.cuda_stream, ref.data_ptr(), ref.data_ptr() + 8 * n_counts)
                ws = whole.finish(stream.cuda_stream, fetch=False).stats
                rec["reduced_tensor_equals_single_gpu_run"] = bool(torch.equal(ref, buf))
                rec["single_gpu_ms_on_rank0"] = ws["kernel_ms"]
                whole.close()
        barrier()
        eng.close()
        if rank == 0:
            if name == "config1" and world == 1:
                rec.update(cli_wall_config1(synth, w))
            out[name] = rec
        del buf
    return out if rank == 0 else None


if __name__ == "__main__":
    sys.exit(main())
