#!/usr/bin/env python3
"""Executed warp instructions and stall samples of one profiled launch, per SOURCE line of the kernel body.

  python tools/profile_by_line.py REPORT.ncu-rep [LIB.so] [--top N] [--by outer|chain]

Joins the SASS view of an `ncu --set full --import-source on` report (per-instruction "Instructions Executed" and
"# Samples", in address order) with the inline chains nvdisasm -gi prints for the same kernel of the library the
report was taken from (instruction i of the one is instruction i of the other; the opcodes are cross-checked).
`--by outer` charges every instruction to the line of the kernel body it was inlined into."""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_lines  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mangled_hint(kernel_name):
    """'void procell_b200::k_proliferate_coop<(int)32, (bool)0, (bool)1, (int)1>(...)' -> 'coopILi32ELb0ELb1ELi1E'"""
    m = re.search(r'k_proliferate_coop<([^>]*)>', kernel_name)
    if not m:
        return None
    out = "coopI"
    for part in m.group(1).split(","):
        t, v = re.match(r'\s*\((\w+)\)(\d+)', part).groups()
        out += ("Li%sE" % v) if t == "int" else ("Lb%sE" % v)
    return out + "E"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("lib", nargs="?", default=os.path.join(ROOT, "cuda_pro_cell_b200", "libprocell_b200.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--by", default="outer", choices=["outer", "chain"])
    a = ap.parse_args()
    src = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    kname = rows[0][1]
    hdr = rows[1]
    i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    prof = [(r[i_src].strip(), int(r[i_ex]), int(r[i_smp])) for r in rows[2:] if len(r) > i_smp]
    hint = mangled_hint(kname)
    lines = sass_lines.disassemble(a.lib)
    body = None
    for name, b in sass_lines.sections(lines):
        if hint and hint in name:
            body = b
    if body is None:
        sys.exit("kernel %s not found in %s" % (hint, a.lib))
    # walk the disassembly: one chain per instruction
    chains, ops = [], []
    chain, pending = [], []
    for l in body:
        m = sass_lines.LINE_RE.search(l)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        mi = sass_lines.INSN_RE.match(l)
        if mi:
            if pending:
                chain, pending = pending, []
            chains.append(tuple(chain))
            ops.append(mi.group(1).split(".")[0])
    if len(chains) != len(prof):
        sys.exit("instruction count mismatch: report %d, library %d (profile taken from another build?)" % (len(prof), len(chains)))
    bad = sum(1 for (s, _, _), op in zip(prof, ops) if op not in s)
    if bad > len(prof) // 50:
        sys.exit("opcodes disagree on %d of %d instructions: not the same build" % (bad, len(prof)))
    ex, smp = collections.Counter(), collections.Counter()
    for ch, (_, e, s) in zip(chains, prof):
        key = (ch[-1] if ch else ("?", 0)) if a.by == "outer" else ch
        ex[key] += e
        smp[key] += s
    tot_e, tot_s = sum(ex.values()), sum(smp.values())
    print("%s\n%d warp instructions executed, %d stall samples" % (kname, tot_e, tot_s))
    print("%8s %6s %6s  line" % ("exec", "%exec", "%smp"))
    for key, e in ex.most_common(a.top):
        label = "%s:%d" % key if a.by == "outer" else " <- ".join("%s:%d" % k for k in key)
        print("%8.3fM %5.1f%% %5.1f%%  %s" % (e / 1e6, 100.0 * e / tot_e, 100.0 * smp[key] / max(tot_s, 1), label))


if __name__ == "__main__":
    main()
