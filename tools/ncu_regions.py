#!/usr/bin/env python3
"""Where do the dynamic warp instructions of a specialised kernel go?  Joins the per-SASS-instruction execution counts of an
.ncu-rep (source page) with the line table of the same kernel (nvdisasm --print-line-info on the object file it was built from)
and sums them per REGION: a generated function of spec_<cfg>.cu (res_<m>, score_<m>, full_<m>, pm2_tail ...), or the kernel
body (mpc_spec.cuh outside the helper functions).  Inlined helpers (sub_u8x4, prmt, ...) count for the region that calls them.

usage: ncu_regions.py REPORT.ncu-rep OBJECT.o GENERATED.cu UNITS      (UNITS = tiles of 32 blocks the launch processed)
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_with_lines(obj):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    out = []
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "--print-line-info", cb], capture_output=True, text=True).stdout
        cur = None
        for l in txt.splitlines():
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                toks = m.group(2).split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                out.append((int(m.group(1), 16), op, cur))
    return out


def main():
    rep, obj, gen, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    bounds = []
    for i, l in enumerate(open(gen), 1):
        m = re.match(r"\s*__device__ (?:static )?__forceinline__ \w+ (\w+)\(", l)
        if m:
            bounds.append((i, m.group(1)))

    def fn(line):
        name = "?"
        for b, n in bounds:
            if line >= b:
                name = n
        return name

    sass = sass_with_lines(obj)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    if len(body) != len(sass):
        print(f"warning: {len(body)} profiled instructions vs {len(sass)} in the object file; joining by position", file=sys.stderr)
    genbase = os.path.basename(gen)
    ctx = "kernel"
    per = collections.OrderedDict()
    ops = collections.defaultdict(collections.Counter)
    smp = collections.Counter()
    for r, (addr, op, loc) in zip(body, sass):
        if loc and loc[0] == genbase:
            ctx = fn(loc[1])
        elif loc and loc[0] == "mpc_spec.cuh" and loc[1] > 330:  # kernel body (helpers live above)
            ctx = "kernel"
        n = float(r[ie] or 0)
        per[ctx] = per.get(ctx, 0.0) + n
        ops[ctx][op.split(".")[0]] += n
        smp[ctx] += float(r[ns] or 0)
    tot = sum(per.values())
    stot = sum(smp.values()) or 1
    print(f"total {tot / units:.1f} warp instructions per tile")
    allops = collections.Counter()
    for c in ops:
        allops.update(ops[c])
    print("  opcodes: " + ", ".join(f"{k} {v / units:.0f}" for k, v in allops.most_common(14)))
    for c, n in per.items():
        top = ", ".join(f"{k} {v / units:.0f}" for k, v in ops[c].most_common(6))
        print(f"  {c:16s} {n / units:8.1f}  samples {smp[c] / stot * 100:5.1f}%   {top}")


if __name__ == "__main__":
    main()
