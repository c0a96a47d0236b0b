#!/usr/bin/env python
"""profiles/*_sass_extract.txt: which Blackwell instructions the built library really contains (cuobjdump -sass, no GPU needed).
Counts of static instructions of each kind, then the first lines of each kind with the kernel they belong to.
    python tools/sass_extract.py > profiles/r2_sass_extract.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "phamers_b200", "lib", "libphamers_b200.so")
KINDS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "ATOMS", "ATOMG", "REDG"]
SHOW = {"ATOMS.POPC.INC": 6, "UBLKCP": 6, "UTMALDG": 6, "LDTM": 4, "UTCHMMA": 6, "UTCBAR": 6}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, first = collections.Counter(), collections.defaultdict(list)
    func = ""
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            func = m.group(1)
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        if not any(op.startswith(k) for k in KINDS):
            continue
        name = ".".join(op.split(".")[:3]) if op.startswith(("ATOMS", "SYNCS", "REDG", "ATOMG", "UTCATOMSWS")) else op.split(".")[0] + ("." if op.startswith("LDTM") else "")
        if op.startswith("UBLKCP") or op.startswith("UTMALDG"):
            name = ".".join(op.split(".")[:3]) if op.startswith("UBLKCP") else ".".join(op.split(".")[:2])
        counts[name] += 1
        for key, n in SHOW.items():
            if op.startswith(key) and len(first[key]) < n:
                first[key].append("%-28s %s" % (func[:28], line.strip()[:150]))
    print("# SASS evidence from phamers_b200/lib/libphamers_b200.so (cuobjdump -sass, %s only): tcgen05 = UTC*MMA / LDTM, TMA = UTMALDG / UBLKCP," % ", ".join(archs))
    print("# mbarrier = SYNCS, shared-memory atomics of the histogram = ATOMS.POPC.INC (ATOMS.ADD: the packed 16-bit table of k = 5).")
    print("# Counts of static instructions, then the first lines of each kind.  Regenerate: python tools/sass_extract.py\n")
    for name, n in counts.most_common():
        print("%8d  %s" % (n, name))
    print()
    for key in SHOW:
        for line in first[key]:
            print(line)


if __name__ == "__main__":
    sys.exit(main())
