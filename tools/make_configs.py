#!/usr/bin/env python3
"""Writes the MPC (VPC) JSON configs shipped in configs/.

The reference ships no config; its schema is what VPC::parseConfig reads
(/root/reference/src/compressor/VPC.cpp:72-330).  Configs written here:

  P6.json   the survey's probe config: AllZero, AllWordSame and four
            PredComp modules (Consecutive, OneBase, DiffBase stride 4, WeightBase) that all use
            the plane-major scan "for plane r: columns 1..127,0".  Known-answer vectors in
            SURVEY.md section 8c are quoted on this config.
  F4.json   delta-friendly config for fp32/int32 arrays: scan visits the
            columns by byte significance (byte 3 of every word, then byte 2, ...), all 8 planes
            of a column together; DiffBase stride 4 / stride 8, OneBase, DiffBase stride 1.
  Z1.json   AllZero + one PredComp only (exercises compressLineOnlyAllZero, VPC.cpp:54-70).
  E5.json   explicit encoding_bits list, 5 modules, mixed scans (plane-major and column-major).
  S32.json  F4's design for 32-byte lines (GPGPU-Sim sector traces, ACCESS_GRAN 32; LoaderGPGPU.cpp) and
  S64.json  for 64-byte lines: the reference takes any lineSize (VPC.cpp:99-101).
"""
import json
import os
import sys

L = 128


def plane_major_scan(cols):
    rows, cs = [], []
    for r in range(8):
        for c in cols:
            rows.append(r)
            cs.append(c)
    return {"TableSize": len(rows), "Rows": rows, "Cols": cs}


def column_major_scan(cols):
    rows, cs = [], []
    for c in cols:
        for r in range(8):
            rows.append(r)
            cs.append(c)
    return {"TableSize": len(rows), "Rows": rows, "Cols": cs}


def fpc_stub():
    # parsed by the reference (VPC.cpp:214-303) and then ignored
    return {"num_modules": 2,
            "0": {"name": "ZerosPattern", "encodingBitsZRLE": 7, "encodingBitsZero": 4},
            "1": {"name": "UncompressedPattern", "encodingBits": 17}}


def predcomp(pred, consecutive_xor, scan):
    return {"name": "PredComp", "submodules": {
        "ResidueModule": {"PredictorModule": pred},
        "XORModule": {"consecutiveXOR": consecutive_xor},
        "ScanModule": scan,
        "FPCModule": fpc_stub()}}


def one_base(root=0):
    return {"name": "OneBasePredictor", "LineSize": L, "RootIndex": root}


def consecutive(root=0):
    return {"name": "ConsecutiveBasePredictor", "LineSize": L, "RootIndex": root}


def diff_base(stride, root=0, diff=None):
    base = [0 if i < stride else i - stride for i in range(L)]
    return {"name": "DiffBasePredictor", "LineSize": L, "RootIndex": root,
            "BaseIndexTable": base, "DiffTable": diff if diff is not None else [0] * L}


def weight_base(root=0):
    base = [0] + [i - 1 for i in range(1, L)]
    w = [1.0] + [[1.0, 0.5, 2.0, 0.3][i % 4] for i in range(1, L)]
    return {"name": "WeightBasePredictor", "LineSize": L, "RootIndex": root,
            "BaseIndexTable": base, "WeightTable": w}


def config(modules, encoding_bits=None):
    ov = {"num_modules": len(modules), "lineSize": L}
    if encoding_bits is not None:
        ov["encoding_bits"] = encoding_bits
    return {"overview": ov, "modules": {str(i): m for i, m in enumerate(modules)}}


def p6():
    scan = plane_major_scan(list(range(1, L)) + [0])
    return config([
        {"name": "AllZero"}, {"name": "AllWordSame"},
        predcomp(consecutive(), True, scan),
        predcomp(one_base(), False, scan),
        predcomp(diff_base(4), True, scan),
        predcomp(weight_base(), False, scan)])


def significance_cols():
    cols = []
    for b in (3, 2, 1, 0):
        cols += [c for c in range(4 + b, L, 4)]
    return cols + [3, 2, 1, 0]


def f4():
    scan = column_major_scan(significance_cols())
    return config([
        {"name": "AllZero"}, {"name": "AllWordSame"},
        predcomp(diff_base(1), True, scan),
        predcomp(one_base(), True, scan),
        predcomp(diff_base(8), True, scan),
        predcomp(diff_base(4), True, scan)])


def z1():
    scan = column_major_scan(list(range(1, L)) + [0])
    return config([{"name": "AllZero"}, predcomp(diff_base(4), False, scan)])


def e5():
    pm = plane_major_scan(list(range(1, L)) + [0])
    cm = column_major_scan(significance_cols())
    diff = [(3 * i) % 7 for i in range(L)]
    return config([
        {"name": "AllZero"}, {"name": "ByteplaneAllSame"},
        predcomp(diff_base(4, diff=diff), True, pm),
        predcomp(one_base(), False, cm),
        predcomp(diff_base(2), True, cm)],
        encoding_bits=[2, 1, 3, 4, 2, 5])


def with_line_size(n, make):
    global L
    keep, L = L, n
    try:
        return make()
    finally:
        L = keep


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "configs")
    os.makedirs(out, exist_ok=True)
    for name, cfg in (("P6", p6()), ("F4", f4()), ("Z1", z1()), ("E5", e5()), ("S32", with_line_size(32, f4)), ("S64", with_line_size(64, f4))):
        with open(os.path.join(out, name + ".json"), "w") as f:
            json.dump(cfg, f, separators=(",", ":"))
            f.write("\n")
    print("wrote P6 F4 Z1 E5 S32 S64 to", out)


if __name__ == "__main__":
    main()
