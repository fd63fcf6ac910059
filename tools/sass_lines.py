#!/usr/bin/env python3
"""Static instruction accounting of one kernel from the built library (no GPU needed).

  python tools/sass_lines.py [LIB.so] [--kernel SUBSTR] [--by outer|inner|chain] [--top N]

Extracts the sm_100a cubin (cuobjdump -xelf), disassembles it with inlining information (nvdisasm -gi) and counts
SASS instructions per source line.  `--by outer` attributes every instruction to the line of the KERNEL body it
was inlined into (so "Philox + ziggurat of the DIVIDE iteration" is one row), `--by inner` to the innermost
line, `--by chain` to the whole inline chain.  The kernels of this repo are issue-bound, so
instruction counts of the straight-line DIVIDE / SEED iterations are the first-order cost model used when tuning
without a GPU at hand."""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE_RE = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
INSN_RE = re.compile(r'^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P(?:\d+|T)\s+)?([A-Z0-9_.]+)')


def disassemble(lib):
    tmp = tempfile.mkdtemp(prefix="sass_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cubins = [f for f in os.listdir(tmp) if f.startswith("sim_kernels.") and f.endswith(".cubin")]
    if not cubins:
        sys.exit("no sim_kernels cubin in " + lib)
    out = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubins[0])], check=True, capture_output=True, text=True)
    return out.stdout.splitlines()


def sections(lines):
    name, start = None, 0
    for i, l in enumerate(lines):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            if name:
                yield name, lines[start:i]
            name, start = m.group(1), i
    if name:
        yield name, lines[start:]


def account(body, by):
    counts = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    chain = []
    pending = []
    for l in body:
        m = LINE_RE.search(l)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        mi = INSN_RE.match(l)
        if mi:
            if pending:
                chain, pending = pending, []
            if not chain:
                key = ("?", 0)
            elif by == "outer":
                key = chain[-1]
            elif by == "inner":
                key = chain[0]
            else:
                key = tuple(chain)
            counts[key] += 1
            ops[key][mi.group(1).split(".")[0]] += 1
    return counts, ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib", nargs="?", default=os.path.join(ROOT, "cuda_pro_cell_b200", "libprocell_b200.so"))
    ap.add_argument("--kernel", default="k_proliferate_coopILi32ELb0ELb1ELi1ELi0E")
    ap.add_argument("--by", default="outer", choices=["outer", "inner", "chain"])
    ap.add_argument("--top", type=int, default=0)
    ap.add_argument("--range", default="", help="only outer lines A-B of sim_kernels.cu, e.g. 504-591")
    a = ap.parse_args()
    lines = disassemble(a.lib)
    for name, body in sections(lines):
        if a.kernel not in name:
            continue
        counts, ops = account(body, a.by)
        total = sum(counts.values())
        print("== %s: %d instructions" % (name, total))
        items = sorted(counts.items(), key=(lambda kv: -kv[1]) if a.top else (lambda kv: kv[0]))
        if a.top:
            items = items[:a.top]
        lo, hi = (map(int, a.range.split("-")) if a.range else (0, 1 << 30))
        sub = 0
        for key, n in items:
            k0 = key if a.by != "chain" else key[-1]
            if a.by != "inner" and not (lo <= k0[1] <= hi):
                continue
            sub += n
            top_ops = " ".join("%s:%d" % kv for kv in ops[key].most_common(6))
            label = "%s:%d" % key if a.by != "chain" else " <- ".join("%s:%d" % k for k in key)
            print("%6d  %-40s %s" % (n, label, top_ops))
        print("subtotal %d" % sub)


if __name__ == "__main__":
    main()
