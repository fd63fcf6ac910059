"""Static checks of the built sm_100a code (no GPU): resource usage and the SASS-level invariants the design relies on.

A 32-warp instance that needs more than 64 registers, or spills, cannot run as one 1024-thread CTA per SM at full
speed; the ring layout is only worth its name if pops and pushes really are 128-bit shared-memory accesses; the
fresh-node path needs its vote.  These regress silently on a box without a GPU unless they are tested."""
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
    assert len(coop) == 19, "CTA shapes (32 / 24 / 16 warps, 16 warps x 2 nodes per lane) x histogram mode x PLAIN + 2 subtree-sharding instances + the set-direct sweep instance"
    for name, u in coop.items():
        warps = int(re.search(r"coopILi(\d+)E", name).group(1))
        plain = re.search(r"coopILi\d+ELb[01]ELb1E", name) is not None
        # PLAIN (one parameter set, one checkpoint: configs 1-4 and the bench) must not touch local memory at all;
        # the general instances (sweeps, time series) are allowed the few words ptxas keeps on the stack today
        # (8-24 bytes, stored in the prologue) - more than that means the 64-register budget no longer holds.
        product = "coopILi32ELb0ELb1ELi1ELi0E" in name     # 32 warps, direct table, PLAIN, MODE 0: configs 1-4 and the bench
        # the product instance must not touch local memory at all; the other PLAIN instances may keep ONE rarely used
        # scalar (the donation epoch, read every 8th iteration) on the stack, the general ones a few words
        assert u["LOCAL"] == 0 and u["STACK"] <= (0 if product else 8 if plain else 24), "%s spills (%d bytes)" % (name, u["STACK"])
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
    assert not re.search(r"\b(LDL|STL)\b", text), "local-memory traffic in the product kernel"
    assert "LDS.128" in text and "STS.128" in text, "ring pops / pushes are no longer 16-byte accesses"
    assert re.search(r"@!?P\d+\s+STS\.128", text), "pushes are no longer predicated stores"
    assert "VOTE.ALL" in text, "the fresh-node vote is gone"
    assert "ATOMS.ADD" in text and "MATCH" not in text, "direct-mode leaf count should be one shared atomic per lane"
    assert "DFMA" in text and "MUFU.RSQ64H" in text
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
