#!/usr/bin/env python3
"""A/B helper for the GPU box: PATTERN (mpc_pattern_run_device: analysis kernel + hash sort + duplicate pass) over synthetic
dumps resident in HBM for one library build (MPC_B200_LIB), with a fingerprint of every counter and histogram bin so that two
builds can be compared.  usage: MPC_B200_LIB=... python tools/ab_pattern.py LABEL [--gib 1] [--line 128] kind ..."""
import argparse
import hashlib
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("label")
    ap.add_argument("kinds", nargs="+")
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--line", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    mpcb = importlib.import_module("cal_22-mpc_b200")
    n128 = int(a.gib * (1 << 30)) // 128
    m = mpcb.Mpc(os.path.join(ROOT, "configs", "F4.json"))
    d = torch.empty(n128 * 128, dtype=torch.uint8, device="cuda")
    for kind in a.kinds:
        m.synth_device(d.data_ptr(), 0, n128, n128, kind, 31337)
        m.sync()
        n = n128 * (128 // a.line)
        best = 1e30
        for _ in range(a.reps):
            _, ps, ms = mpcb.pattern_run(device_ptr=d.data_ptr(), n_blocks=n, line_size=a.line)
            best = min(best, ms)
        h = hashlib.sha256()
        for f in ("symbol_counts", "symbol_counts_nontrivial", "implicit_bytes", "explicit_bytes"):
            h.update(np.asarray(list(getattr(ps, f)), dtype=np.uint64).tobytes())
        for f in ("blocks", "zeros_bytes", "repeated_bytes", "undefined_bytes", "temporal_bytes", "distinct_blocks"):
            h.update(int(getattr(ps, f)).to_bytes(8, "little"))
        print(f"{a.label:8s} L={a.line:3d} {kind:14s} {best:8.3f} ms {n128 * 128 / best / 1e6:8.1f} GB/s  path {ps.temporal_path}  fp {h.hexdigest()[:16]}", flush=True)


if __name__ == "__main__":
    main()
