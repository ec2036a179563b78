"""Static evidence for the built library (no GPU needed): per kernel, the SASS mnemonics that prove the Blackwell
paths (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, HMMA = legacy mma.sync)
plus registers / shared memory / spills from `cuobjdump -res-usage`.

    python tools/sass_evidence.py > profiles/rNN_sass_evidence.md
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "sylph_few_shot_detection_b200", "libsylph_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS",
             "ACQBULK", "PREEXIT"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = name.replace("sylph::", "")
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)
    name = name.replace("(bool)0", "false").replace("(bool)1", "true").replace("(int)", "")
    return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            kv = dict(re.findall(r"(\w+):(\d+)", line))
            usage[cur] = kv
            cur = None
    counts = {}
    cur = None
    ninstr = {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {k: 0 for k in MNEMONICS}
            ninstr[cur] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        ninstr[cur] += 1
        op = m.group(1)
        for k in MNEMONICS:
            if op.startswith(k):
                counts[cur][k] += 1
    names = demangle(list(counts))
    rows = []
    for k, c in counts.items():
        u = usage.get(k, {})
        rows.append((short(names[k]), ninstr[k], u.get("REG", "?"), u.get("SHARED", "?"), u.get("STACK", "?"),
                     u.get("LOCAL", "?"), c))
    rows.sort(key=lambda r: (-(r[6]["UTCHMMA"] + r[6]["UTCQMMA"]), -(r[6]["UTMALDG"] + r[6]["UTMASTG"]), r[0]))
    print("# SASS / resource evidence of `libsylph_b200.so` (static; `python tools/sass_evidence.py`)\n")
    print("Built with `nvcc " + " ".join(__import__("sylph_few_shot_detection_b200._lib", fromlist=["x"]).NVCC_FLAGS) + "`.")
    print("Mnemonics as in `/opt/skills/guides/B200_PROFILING.md`: `UTCHMMA` = `tcgen05.mma kind::f16`, `LDTM` = `tcgen05.ld`, "
          "`UTMALDG`/`UTMASTG` = TMA tensor load / store, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier ops, "
          "`ACQBULK`/`PREEXIT` = programmatic dependent launch (`griddepcontrol`), `HMMA` = legacy `mma.sync` (must be 0).\n")
    tensor = [r for r in rows if r[6]["UTCHMMA"] + r[6]["UTCQMMA"] > 0]
    other = [r for r in rows if r not in tensor]
    hdr = "| kernel | SASS instr | regs | static smem B | stack B | UTCHMMA | LDTM | UTMALDG | UTMASTG | UTCBAR | SYNCS | HMMA |\n|---|---|---|---|---|---|---|---|---|---|---|---|"
    print(f"## Tensor-core kernels ({len(tensor)} instantiations)\n")
    print(hdr)
    for n, ni, reg, sh, st, lo, c in tensor:
        print(f"| `{n}` | {ni} | {reg} | {sh} | {st} | {c['UTCHMMA'] + c['UTCQMMA']} | {c['LDTM']} | {c['UTMALDG']} | "
              f"{c['UTMASTG']} | {c['UTCBAR']} | {c['SYNCS']} | {c['HMMA']} |")
    print(f"\n## Memory-bound / helper kernels ({len(other)})\n")
    print("| kernel | SASS instr | regs | static smem B | stack B | UTMALDG | UTMASTG | LDGSTS | HMMA |\n|---|---|---|---|---|---|---|---|---|")
    for n, ni, reg, sh, st, lo, c in other:
        print(f"| `{n}` | {ni} | {reg} | {sh} | {st} | {c['UTMALDG']} | {c['UTMASTG']} | {c['LDGSTS']} | {c['HMMA']} |")
    spills = [(n, st) for n, ni, reg, sh, st, lo, c in rows if st not in ("0", "?")]
    print("\n## Stack frames (local arrays or spills; `-Xptxas -v` tells which)\n")
    if spills:
        for n, st in spills:
            print(f"* `{n}`: {st} B of stack")
    else:
        print("No kernel uses stack.")
    tot = {k: sum(r[6][k] for r in rows) for k in MNEMONICS}
    print("\nTotals over the library: " + ", ".join(f"{k} {v}" for k, v in tot.items()) + ".")


if __name__ == "__main__":
    sys.path.insert(0, REPO)
    main()
