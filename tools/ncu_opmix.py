#!/usr/bin/env python3
"""Dynamic SASS opcode mix of one profiled kernel: ncu_opmix.py REPORT.ncu-rep [units]  (units = blocks/32 tiles to normalise by)"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ci = {k: i for i, k in enumerate(h)}
    tot = 0
    byop = collections.Counter()
    samples = collections.Counter()
    nlines = 0
    for r in rows[2:]:
        if len(r) != len(h):
            continue
        n = float(r[ci["Instructions Executed"]] or 0)
        sass = r[ci["Source"]].split()
        op = sass[1] if sass[0].startswith("@") else sass[0]
        op = op.split(".")[0]
        byop[op] += n
        samples[op] += float(r[ci["# Samples"]] or 0)
        tot += n
        nlines += 1
    print(f"static SASS lines {nlines} ({nlines * 16 / 1024:.1f} KiB), dynamic warp instructions {tot:.0f}")
    stot = sum(samples.values())
    for op, n in byop.most_common(30):
        extra = f"  {n / units:8.1f}/tile" if units else ""
        print(f"{op:12s} {n / tot * 100:6.2f}%{extra}   samples {samples[op] / stot * 100:5.1f}%")


if __name__ == "__main__":
    main()
