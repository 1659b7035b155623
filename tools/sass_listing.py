#!/usr/bin/env python3
"""SASS evidence for profiles/: for every specialised kernel in libmpc_b200.so, the opcode histogram of the whole kernel
and the listing of its hot region (from the tile read LDS.128s to the histogram atomic), straight from `cuobjdump -sass`.

  python tools/sass_listing.py [--out profiles] [--full]     (--full keeps every line instead of the head of each kernel)
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cal_22-mpc_b200", "libmpc_b200.so")


def kernels(sass):
    cur, name = [], None
    for ln in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            cur.append(ln.rstrip())
    if name:
        yield name, cur


def opcode(ln):
    body = ln.split("*/", 1)[1].strip()
    toks = body.split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    return toks[0].rstrip(";") if toks else "?"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles"))
    ap.add_argument("--head", type=int, default=400, help="instructions of the hot region to keep per kernel")
    ap.add_argument("--prefix", default="sass_")
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
    for name, lines in kernels(sass):
        dn = demangle(name)
        m = re.search(r"spec_(\w+)::Cfg", dn)
        if not m:
            continue
        cfg = m.group(1)
        hist = collections.Counter(opcode(l) for l in lines)
        # hot region: from the first 128-bit shared load of the tile to the histogram atomic
        first = next((i for i, l in enumerate(lines) if "LDS.128" in l), 0)
        last = max((i for i, l in enumerate(lines) if "ATOMS" in l), default=len(lines) - 1)
        path = os.path.join(a.out, f"{a.prefix}spec_{cfg}.txt")
        with open(path, "w") as f:
            f.write(f"# cuobjdump -sass cal_22-mpc_b200/libmpc_b200.so -- {dn}\n")
            f.write(f"# {len(lines)} SASS instructions in the kernel; tile loop = instructions {first}..{last} ({last - first + 1})\n")
            f.write("# opcode histogram of the whole kernel (static counts):\n")
            for op, c in hist.most_common(40):
                f.write(f"#   {c:6d}  {op}\n")
            f.write(f"# --- first {a.head} instructions of the tile loop (tile read, zero / word-repeat test, scoring) ---\n")
            for l in lines[first:first + a.head]:
                f.write(l + "\n")
            f.write(f"# --- last {a.head // 2} instructions of the tile loop (end of the encoder, statistics, per-block store) ---\n")
            for l in lines[max(first, last - a.head // 2):last + 8]:
                f.write(l + "\n")
        print(path, len(lines), "instructions", file=sys.stderr)


if __name__ == "__main__":
    main()
