#!/usr/bin/env python
"""Compare the SASS of two builds of libsylph_b200.so kernel by kernel (no GPU needed): which kernels are byte-identical
in their instruction streams, which changed, which are new.  Used to prove that an edit behind `if constexpr` or a new
opt-in kernel leaves every kernel of the measured default path untouched.

    cp sylph_few_shot_detection_b200/libsylph_b200.so /tmp/before.so   # then edit + rebuild
    python tools/sass_diff.py /tmp/before.so sylph_few_shot_detection_b200/libsylph_b200.so

Template kernels are matched after demangling, ignoring trailing defaulted template arguments that were added."""
import re
import subprocess
import sys


def kernels(path):
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    ks, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            ks[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if cur and m:
            ks[cur].append(m.group(1).strip())
    names = subprocess.run(["c++filt"], input="\n".join(ks), capture_output=True, text=True).stdout.splitlines()
    out = {}
    for mangled, name in zip(ks, names):
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)
        out[name] = ks[mangled]
    return out


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])

    def match(name):
        if name in b:
            return name
        stem = name[:-1] if name.endswith(">") else name       # "k<1, 2>" matches "k<1, 2, false>"
        cands = [n for n in b if n.startswith(stem + ",") and n not in a]
        same = [n for n in cands if b[n] == a[name]]
        return (same or cands or [None])[0]

    identical, changed, gone, used = [], [], [], set()
    for name in sorted(a):
        m = match(name)
        if m is None:
            gone.append(name)
        else:
            used.add(m)
            (identical if a[name] == b[m] else changed).append((name, m))
    new = sorted(n for n in b if n not in used)
    print(f"{len(identical)} kernels identical, {len(changed)} changed, {len(gone)} removed, {len(new)} new")
    for name, m in changed:
        print(f"  CHANGED  {name}  ({len(a[name])} -> {len(b[m])} instructions)")
    for name in gone:
        print(f"  REMOVED  {name}")
    for name in new:
        print(f"  NEW      {name}  ({len(b[name])} instructions)")
    return 1 if changed or gone else 0


if __name__ == "__main__":
    sys.exit(main())
