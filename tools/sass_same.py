#!/usr/bin/env python3
"""Do two builds of the library contain the same machine code for a kernel?  (no GPU needed)

  python tools/sass_same.py OLD.so NEW.so [KERNEL_SUBSTR ...]      # default: every kernel of sim_kernels.cu

Compares the SASS instruction streams with branch labels normalised.  Used to show that a refactor or a comment edit
left a GPU-verified kernel untouched (line-number tables change, instructions must not)."""
import difflib
import os
import re
import subprocess
import sys
import tempfile


def kernels(lib):
    tmp = tempfile.mkdtemp(prefix="sass_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith("sim_kernels.") and f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cub)], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for l in out.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            # k_proliferate_coop gained a fifth template argument (MODE); <..., 0> is the kernel it was before
            name = re.sub(r'(k_proliferate_coopILi\d+ELb[01]ELb[01]ELi\d+E)L[bi]0E(EEvNS_9SimParamsE)', r'\1\2', m.group(1))
            res[name] = []
            continue
        mi = re.match(r'^\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
        if mi and name:
            insn = re.sub(r'`\(\.L_x_\d+\)', 'L', mi.group(1)).strip()
            # names of anonymous-namespace callees carry a hash of the source PATH: the same code built in another
            # directory must still compare equal
            insn = re.sub(r'(_INTERNAL_|_GLOBAL__N__)[0-9a-f]{8}_', r'\1', insn)
            insn = re.sub(r'\$__internal_\d+_\$', '$__internal_$', insn)      # libdevice helpers are numbered per file
            res[name].append(re.sub(r'(k_proliferate_coopILi\d+ELb[01]ELb[01]ELi\d+E)L[bi]0E(EEvNS_9SimParamsE)', r'\1\2', insn))
    return res


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    want = sys.argv[3:]
    bad = 0
    for name in sorted(set(a) | set(b)):
        if want and not any(w in name for w in want):
            continue
        x, y = a.get(name), b.get(name)
        if x is None or y is None:
            print("ONLY IN ONE  %s" % name)
            bad += 1
            continue
        d = [l for l in difflib.unified_diff(x, y, lineterm='', n=0) if l[:1] in "+-" and l[:3] not in ("+++", "---")]
        print("%-9s %5d instructions  %s" % ("same" if not d else "DIFFERENT", len(y), name))
        bad += bool(d)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
