#!/usr/bin/env python3
"""A/B helper for the GPU box: device-timed GB/s of one library build (MPC_B200_LIB) over a list of (config, kind)
workloads resident in HBM, with the compressed total of each as a fingerprint (equal fingerprints across builds = same
statistics).  usage: MPC_B200_LIB=... python tools/ab_bench.py LABEL [--gib 1] [--reps 20] [--packed] cfg:kind ..."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("label")
    ap.add_argument("work", nargs="+")
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--packed", action="store_true")
    a = ap.parse_args()
    mpcb = importlib.import_module("cal_22-mpc_b200")
    n = int(a.gib * (1 << 30)) // 128
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    packed = torch.zeros(n * 4, dtype=torch.int16, device="cuda")
    ctxs, last_kind = {}, None
    out = []
    nbytes = n * 128
    synth_ctx = mpcb.Mpc(os.path.join(ROOT, "configs", "F4.json"))  # synthetic dumps are defined on 128-byte blocks
    synth_ctx.set_stream(stream.cuda_stream)
    for w in a.work:
        cfg, kind = w.split(":")[:2]
        generic = w.endswith(":generic")  # cfg:kind:generic forces the warp-per-block kernel
        if (cfg, generic) not in ctxs:
            ctxs[(cfg, generic)] = mpcb.Mpc(os.path.join(ROOT, "configs", cfg + ".json"))
            ctxs[(cfg, generic)].set_stream(stream.cuda_stream)
            if generic:
                ctxs[(cfg, generic)].set_kernel(1)
        m = ctxs[(cfg, generic)]
        n = nbytes // m.line_size  # lines of this config's size over the same bytes
        if kind != last_kind:
            synth_ctx.synth_device(d.data_ptr(), 0, nbytes // 128, nbytes // 128, kind, 31337)
            last_kind = kind
        pp = packed.data_ptr() if a.packed else None
        for _ in range(3):
            m.submit_device(d.data_ptr(), n, pp)
        torch.cuda.synchronize()
        m.reset()
        m.enable_timing(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.reps):
            m.submit_device(d.data_ptr(), n, pp)
        e1.record(stream)
        torch.cuda.synchronize()
        m.enable_timing(True)
        ms = e0.elapsed_time(e1) / a.reps
        st = m.finish()
        out.append(f"{w}={nbytes / ms / 1e6:.0f}")
        print(f"AB {a.label:14s} {w:22s} {nbytes / ms / 1e6:8.1f} GB/s  {ms:7.4f} ms  comp_bits/rep {st.CompressedSize // a.reps}  res_sq {int(st.res_sq.sum()) // a.reps}", flush=True)
    print("ABSUM", a.label, " ".join(out), flush=True)


if __name__ == "__main__":
    main()
