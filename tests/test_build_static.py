"""Static checks of the built sm_100a code (no GPU): resource usage and the SASS-level invariants the design relies on.

A 32-warp instance that needs more than 64 registers, or spills, cannot run as one 1024-thread CTA per SM at full
speed; the ring layout is only worth its name if pops and pushes really are 128-bit shared-memory accesses; the
common iteration must stay free of local-memory traffic.  These regress silently on a box without a GPU unless they are tested."""
import re
import shutil
import subprocess

import pytest

from cuda_pro_cell_b200 import _lib

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvdisasm") is None,
                                reason="CUDA binary utilities not installed")


def _res_usage():
    out = subprocess.run(["cuobjdump", "-res-usage", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    usage, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        if name and "REG:" in line:
            usage[name] = {k: int(v) for k, v in re.findall(r"(\w+)(?:\[\d+\])?:(\d+)", line)}
            name = None
    return usage


def test_kernel_instances_fit_their_cta_shape():
    usage = _res_usage()
    coop = {k: v for k, v in usage.items() if "k_proliferate_coop" in k}
    assert len(coop) == 16, ("CTA shapes (32 / 24 / 16 warps) x histogram mode x PLAIN + 2 subtree-sharding instances + the set-direct sweep "
                             "instance + the deep-tree instance with merged leaf counts")
    for name, u in coop.items():
        warps = int(re.search(r"coopILi(\d+)E", name).group(1))
        plain = re.search(r"coopILi\d+ELb[01]ELb1E", name) is not None
        # A few words of stack are allowed: ptxas parks rarely used scalars there (the bin hint of the SEED iteration, the
        # division-count correction, values that live across the general DIVIDE iteration).  What must NOT happen is
        # local-memory traffic in the common DIVIDE iteration: test_product_instance_keeps_its_sass_level_shape looks at
        # that block instruction by instruction.  More stack than this means the 64-register budget no longer holds.
        assert u["STACK"] <= (32 if plain else 128), "%s spills (%d bytes)" % (name, u["STACK"])
        assert u["REG"] * warps * 32 <= 65536, "%s: %d registers do not fit %d warps on one SM" % (name, u["REG"], warps)
    assert any("k_rng_ceiling" in k for k in usage) and any("k_proliferate_simple" in k for k in usage)


def test_product_instance_keeps_its_sass_level_shape():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import sass_lines
    lines = sass_lines.disassemble(str(_lib.LIB_PATH))
    body = None
    for name, b in sass_lines.sections(lines):
        if "k_proliferate_coopILi32ELb0ELb1ELi1ELi0E" in name:  # 32 warps, direct histogram, PLAIN, MODE 0: config 2's kernel
            body = [l for l in b if sass_lines.INSN_RE.match(l)]
    assert body, "product instance not found in the library"
    text = "\n".join(body)
    assert "LDS.128" in text and "STS.128" in text, "ring pops / pushes are no longer 16-byte accesses"
    assert re.search(r"@!?P\d+\s+STS\.128", text), "pushes are no longer predicated stores"
    assert "ATOMS.ADD" in text and "MATCH" not in text, "direct-mode leaf count should be one shared atomic per lane"
    assert "DFMA" in text and "MUFU.RSQ64H" in text
    # the common DIVIDE iteration = the straight-line code from a ring pop (two LDS.128 in a row) through exactly one
    # Philox block (20 IMAD.WIDE) to the leaf count (ATOMS.ADD); of the candidates the shortest is the FULL instance (no
    # "lane has no node" branches).  It must not touch local memory and must stay near its instruction budget.
    ops = [sass_lines.INSN_RE.match(l).group(1) for l in body]
    best = None
    for end in (i for i, o in enumerate(ops) if o == "ATOMS.ADD"):
        pops = [i for i in range(max(0, end - 400), end) if ops[i] == "LDS.128" and ops[i + 1] == "LDS.128"
                or (ops[i] == "LDS.128" and i + 2 < len(ops) and ops[i + 2] == "LDS.128")]
        for start in reversed(pops):
            block = ops[start:end + 1]
            if sum(o.startswith("IMAD.WIDE") for o in block) == 20:
                if best is None or len(block) < len(best):
                    best = block
                break
    assert best is not None
    assert not any(o in ("LDL", "STL") or o.startswith("LDL.") or o.startswith("STL.") for o in best), "local-memory traffic in the common DIVIDE iteration"
    n_fp64 = sum(o[0] == "D" and o.split(".")[0] in ("DFMA", "DADD", "DMUL", "DSETP") for o in best)
    assert n_fp64 <= 16 and len(best) <= 165, (n_fp64, len(best))      # 153 in the final build of round 2 (180 one build earlier)
    # the deep-tree instance (MODE 3) is the same kernel with equal leaf keys merged before the atomic
    merged = [b for name, b in sass_lines.sections(lines) if "k_proliferate_coopILi32ELb0ELb1ELi1ELi3E" in name]
    assert merged and any("MATCH.ANY" in l for l in merged[0]), "merged-leaf-count instance missing or without MATCH.ANY"
    counts, _ = sass_lines.account([l for l in lines], "outer")       # whole file: only a smoke test of the tool
    assert sum(counts.values()) > 10000


def test_setdirect_rendezvous_protocol_model(tmp_path):
    """The CTA-level protocol of kernel MODE 2 (sweeps with a set-relative table: claim, park, last warp drains /
    fetches / re-bases / releases) as a host program with threads for warps: it must terminate with exactly the
    expected tensor for several CTA / warp / set / unit shapes (tests/models/setdirect_protocol_model.cpp)."""
    from pathlib import Path
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    src = Path(__file__).resolve().parent / "models" / "setdirect_protocol_model.cpp"
    exe = tmp_path / "model"
    subprocess.run([cxx, "-std=c++17", "-O2", "-pthread", "-o", str(exe), str(src)], check=True)
    for shape in (("4", "8", "37", "23"), ("2", "32", "50", "100"), ("8", "4", "5", "3"), ("3", "5", "1", "1"), ("5", "16", "64", "29")):
        for _ in range(3):
            r = subprocess.run([str(exe), *shape], capture_output=True, text=True, timeout=120)
            assert r.returncode == 0, (shape, r.stdout)
