#!/usr/bin/env python3
"""Hardware-unit fractions of profiled launches, as JSON for bench.py's `roofline.hardware`.

  python tools/hw_fractions.py OUT.json NAME=REPORT.ncu-rep[#k][:divisions[:draws]] ...

Every REPORT is one kernel launch captured with `ncu --set full --clock-control none`.  Per launch: duration, executed
warp instructions, issue-slot fraction (instructions / SMSP cycles elapsed: one issue slot per SMSP and cycle), FP64 /
ALU / FMA-heavy pipe fractions, shared-memory wavefront fraction, shared-atomic wavefronts per second and their share
of the shared-memory pipe, bank conflicts, DRAM bytes.  With the launch's division count the instructions per 32
divisions (one warp-iteration) follow.  The numbers are taken under the profiler (clocks a few percent below a normal
run, caches cold); bench.py quotes them as fractions, never as times."""
import csv
import json
import subprocess
import sys


def raw(rep):
    """REPORT or REPORT#k: the k-th profiled launch of a report that holds several (0 = first)"""
    k = 0
    if "#" in rep:
        rep, k = rep.rsplit("#", 1)
        k = int(k)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + k]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}
    m = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            m[h] = float(v.replace(",", "")) * (scale.get(u, 1.0) if (h.startswith("dram__bytes") or h == "gpu__time_duration.sum") else 1.0)
        except ValueError:
            pass
    m["__kernel"] = rows[2 + k][hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    return m


def pct(m, key):
    return None if key not in m else m[key] / 100.0


def main():
    out_path = sys.argv[1]
    res = {"how": "ncu --set full --clock-control none, one launch each (tools/gpu_visit.sh); fractions of the hardware peak "
                  "ncu reports for that unit, under the profiler"}
    for arg in sys.argv[2:]:
        name, rest = arg.split("=", 1)
        parts = rest.split(":")
        rep = parts[0]
        divisions = float(parts[1]) if len(parts) > 1 and parts[1] else None
        draws = float(parts[2]) if len(parts) > 2 and parts[2] else None
        m = raw(rep)
        inst = m.get("smsp__inst_executed.sum")
        cyc = m.get("smsp__cycles_elapsed.sum")
        t_ms = m.get("gpu__time_duration.sum")
        atom_wf = m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum")
        rec = {"kernel": m["__kernel"][:120], "report": rep.split("/")[-1], "time_ms_under_ncu": t_ms,
               "warp_instructions": inst,
               "issue_slot_frac": inst / cyc if inst and cyc else None,
               "issue_active_frac": pct(m, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "fp64_pipe_frac": pct(m, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "alu_pipe_frac": pct(m, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
               "fma_heavy_pipe_frac": pct(m, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
               "xu_pipe_frac": pct(m, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
               "lsu_pipe_frac": pct(m, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
               "shared_wavefront_frac": pct(m, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
               "shared_atomic_wavefront_frac": pct(m, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed"),
               "shared_atomic_wavefronts_per_s": atom_wf / (t_ms * 1e-3) if atom_wf and t_ms else None,
               "shared_atomic_instructions": m.get("smsp__inst_executed_op_shared_atom.sum"),
               "shared_bank_conflicts": m.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
               "achieved_warps_per_sm": m.get("sm__warps_active.avg.per_cycle_active"),
               "registers_per_thread": m.get("launch__registers_per_thread"),
               "dram_bytes_per_launch": None if "dram__bytes_read.sum" not in m else
               m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
               "dram_frac": pct(m, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
        if divisions and inst:
            rec["divisions"] = divisions
            rec["warp_instructions_per_32_divisions"] = inst / divisions * 32.0
        if draws and inst:
            rec["draws"] = draws
            rec["warp_instructions_per_32_draws"] = inst / draws * 32.0
        res[name] = rec
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
