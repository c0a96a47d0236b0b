#!/usr/bin/env python
"""Per-SASS-instruction attribution of an ncu capture taken with --import-source on (read here, no GPU needed):
shared-memory wavefronts, global sectors and executed warp instructions by opcode, plus the heaviest instructions.

    python tools/ncu_source_summary.py gpurun_out/prof_hist_r2a.ncu-rep [steps] > profiles/r2a_hist_sass_wavefronts.txt

`steps` (optional): number of 512-byte warp-steps of the launch, to print every figure per step as well."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    lines = raw.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    print(lines[0][:200])
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    by_op = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
    tot = [0, 0, 0, 0, 0, 0]
    heavy = []
    for r in rows:
        src = r["Source"].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = op.rstrip(";")
        base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("ATOMS", "LDG", "STG", "LDS", "STS", "SHFL", "RED", "ATOMG") else op.split(".")[0]
        vals = [int(r["Instructions Executed"] or 0), int(r["L1 Wavefronts Shared"] or 0), int(r["L1 Wavefronts Shared Ideal"] or 0),
                int(r["L2 Theoretical Sectors Global"] or 0), int(r["L1 Tag Requests Global"] or 0), int(r["# Samples"] or 0)]
        for i, v in enumerate(vals):
            by_op[base][i] += v
            tot[i] += v
        heavy.append((vals[1] + vals[4], vals, src))
    def fmt(v):
        return "%14d" % v + ("  %8.3f" % (v / steps) if steps else "")
    hdr = "%-14s %s %s %s %s %s %s" % ("opcode", "warp-instr".rjust(14 + (10 if steps else 0)), "smem-wavefr".rjust(14 + (10 if steps else 0)),
                                       "smem-ideal".rjust(14 + (10 if steps else 0)), "l2-sectors".rjust(14 + (10 if steps else 0)),
                                       "l1-tag-req".rjust(14 + (10 if steps else 0)), "samples".rjust(14 + (10 if steps else 0)))
    print(hdr + ("   (second figure of each pair: per 512-byte warp-step)" if steps else ""))
    for op, v in sorted(by_op.items(), key=lambda kv: -(kv[1][1] + kv[1][4] + kv[1][0] * 1e-6)):
        if v[0] == 0:
            continue
        print("%-14s %s %s %s %s %s %s" % (op, fmt(v[0]), fmt(v[1]), fmt(v[2]), fmt(v[3]), fmt(v[4]), fmt(v[5])))
    print("%-14s %s %s %s %s %s %s" % ("TOTAL", fmt(tot[0]), fmt(tot[1]), fmt(tot[2]), fmt(tot[3]), fmt(tot[4]), fmt(tot[5])))
    print("\nheaviest memory instructions (shared wavefronts + global tag requests):")
    for _, vals, src in sorted(heavy, key=lambda t: -t[0])[:40]:
        print("  %12d wavefr %12d ideal %12d tagreq %12d exec   %s" % (vals[1], vals[2], vals[4], vals[0], src[:90]))


if __name__ == "__main__":
    main()
