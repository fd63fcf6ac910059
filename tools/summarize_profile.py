#!/usr/bin/env python3
"""Turn an .ncu-rep (one kernel launch, --set full) into a small tracked text summary under profiles/."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_atom.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
lines = ["# %s" % title, "", "source: `%s` (ncu --set full --clock-control none, one launch)" % rep, "", "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in m:
        lines.append("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
lines += ["", "## warp stall reasons (warps per issue-active cycle)", "", "| reason | value |", "|---|---|"]
st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v.replace(",", "")))
      for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for name, v in sorted(st, key=lambda x: -x[1]):
    lines.append("| %s | %.3f |" % (name, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
if len(srows) > 2:
    h = srows[1]
    ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    body = srows[2:]
    tot = sum(int(r[isamp]) for r in body) or 1
    lines += ["", "## hottest SASS instructions by stall samples (of %d)" % tot, "", "| # | SASS | samples | executed |", "|---|---|---|---|"]
    for i in sorted(sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:25]):
        lines.append("| %d | `%s` | %s | %s |" % (i, body[i][isrc].strip()[:70], body[i][isamp], body[i][ia]))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
