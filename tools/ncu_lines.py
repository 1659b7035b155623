#!/usr/bin/env python3
"""Dynamic warp instructions per CUDA source line of one profiled kernel (needs -lineinfo + --import-source on).
usage: ncu_lines.py REPORT.ncu-rep UNITS [MIN]   (UNITS = tiles to normalise by; lines below MIN instr/unit are folded)"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    lo = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = None, None
    per = collections.OrderedDict()
    samples = collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            ns = hdr.index("# Samples")
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[0]:  # a CUDA source line row: carries the aggregated numbers of its SASS
            key = (fname, int(r[0]), r[1].strip()[:110])
            try:
                per[key] = per.get(key, 0.0) + float(r[ie] or 0)
                samples[key] += float(r[ns] or 0)
            except ValueError:
                pass
    tot = sum(per.values())
    stot = sum(samples.values()) or 1
    print(f"total {tot / units:.1f} warp instructions per unit")
    byfile = collections.Counter()
    for (f, ln, src), n in per.items():
        byfile[f] += n
    for f, n in byfile.most_common():
        print(f"  {f:28s} {n / units:8.1f}")
    for (f, ln, src), n in per.items():
        if n / units >= lo:
            print(f"{f:18s}:{ln:4d} {n / units:7.1f}  smp {samples[(f, ln, src)] / stot * 100:4.1f}%  {src}")


if __name__ == "__main__":
    main()
