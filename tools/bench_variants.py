#!/usr/bin/env python3
"""BASELINE.json config #5: every compressor variant over the same synthetic mixed dump -- throughput (device-timed,
data resident in HBM) and compression ratio next to MPC.  CPACK is sequential by construction and runs on the host.

usage: bench_variants.py [--gib 4] [--kind mixed_hashed] [--config P6]"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=4.0)
    ap.add_argument("--kind", default="mixed_hashed")
    ap.add_argument("--config", default="P6")
    ap.add_argument("--cpack-mib", type=int, default=256)
    ap.add_argument("--only", default=None, help="run just this algorithm (BDI, FPC, BPC, SC2, PATTERN): profiling helper")
    a = ap.parse_args()
    mpcb = importlib.import_module("cal_22-mpc_b200")
    n = int(a.gib * (1 << 30)) // 128
    m = mpcb.Mpc(os.path.join(ROOT, "configs", a.config + ".json"))
    d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    m.synth_device(d.data_ptr(), 0, n, n, a.kind, 31337)
    m.sync()
    rows = []
    for rep in range(0 if a.only else 2):
        m.reset()
        m.submit_device(d.data_ptr(), n, None)
        st = m.finish()
        ms, _ = m.last_timing()
    if not a.only:
        rows.append({"alg": f"MPC ({a.config}, {m.kernel_name()})", "gbs": n * 128 / ms / 1e6, "ratio": st.CompRatio, "where": "GPU"})
    for alg in ("BDI", "FPC", "BPC"):
        if a.only and a.only != alg:
            continue
        for rep in range(2):
            _, vs, ms = mpcb.variant_run(alg, device_ptr=d.data_ptr(), n_blocks=n)
        rows.append({"alg": alg, "gbs": n * 128 / ms / 1e6, "ratio": vs.original_bits / vs.compressed_bits, "where": "GPU"})
    import ctypes as C
    S = mpcb.sc2_sampling_lines(n + 1)
    vs, msf = mpcb.VariantStats(), C.c_float()
    if not a.only or a.only == "SC2":
        for rep in range(2):
            rc = mpcb.lib().mpc_sc2_run_device(0, d.data_ptr(), n, 128, S, None, C.byref(vs), C.byref(msf))
            assert rc == 0, mpcb.lib().mpc_sc2_error()
        rows.append({"alg": f"SC2 (S={S}, incl. sort + host tree)", "gbs": n * 128 / msf.value / 1e6, "ratio": vs.original_bits / vs.compressed_bits, "where": "GPU+host tree"})
    if not a.only or a.only == "PATTERN":
        for rep in range(2):
            _, ps, ms = mpcb.pattern_run(device_ptr=d.data_ptr(), n_blocks=n)
        rows.append({"alg": f"PATTERN (analysis; {ps.distinct_blocks} distinct lines, temporal path {ps.temporal_path})", "gbs": n * 128 / ms / 1e6,
                     "ratio": 0.0, "where": "GPU (kernel + hash sort)"})
    if a.only:
        print(json.dumps(rows))
        return
    nc = min(n, (a.cpack_mib << 20) // 128)
    host = d[: nc * 128].cpu().numpy()
    t0 = time.perf_counter()
    _, cs = mpcb.cpack_run(host)
    dt = time.perf_counter() - t0
    rows.append({"alg": f"CPACK (first {nc * 128 >> 20} MiB, sequential)", "gbs": nc * 128 / dt / 1e9, "ratio": cs.original_bits / cs.compressed_bits, "where": "host, 1 thread"})
    print(f"# {a.gib:g} GiB '{a.kind}' dump, {n} blocks of 128 B")
    for r in rows:
        print(f"{r['alg']:50s} {r['gbs']:10.2f} GB/s   ratio {r['ratio']:.6f}   {r['where']}")
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
