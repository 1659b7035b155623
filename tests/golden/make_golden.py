#!/usr/bin/env python3
"""Generates tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh from
/root/reference).  Run in the build container only -- the GPU box has no /root/reference, it uses the
committed fixtures.

  tests/golden/mpc_blocks.npz    640 blocks (9 known-answer blocks of SURVEY.md section 8c, 8 synthetic classes,
                                 hand-made edge blocks); per shipped config: per-block size + selected
                                 cluster, totals, per-cluster stats, MAE/MSE doubles, histograms -- all
                                 produced by comp::VPC::CompressLine / VPCResult of the reference.
  tests/golden/cli_<cfg>_*.csv   the two CSV files the reference CLI appends for a 200-row .npy (run twice, so
                                 header + 2 rows), plus its stdout line.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bridge import REF_BIN, RefCompressor  # noqa: E402
from tools.gen_dump import kat_blocks, synth, write_gpgpusim_log  # noqa: E402

CONFIGS = ["P6", "F4", "Z1", "E5"]


def edge_blocks():
    rng = np.random.default_rng(99)
    out = []
    for pos in (0, 1, 3, 4, 64, 127):  # a single non-zero byte
        b = np.zeros(128, np.uint8)
        b[pos] = 1 + pos
        out.append(b)
    for v in (1, 0x80, 0xFF):  # all bytes equal
        out.append(np.full(128, v, np.uint8))
    b = np.tile(np.array([1, 2, 3, 4], np.uint8), 32)  # word-same, then broken in the last byte
    out.append(b.copy())
    b[127] ^= 1
    out.append(b.copy())
    b = np.tile(np.array([1, 2, 3, 4], np.uint8), 32)
    b[4] ^= 0x10
    out.append(b)
    for step in (1, 2, 3, 255):  # byte ramps
        out.append((np.arange(128) * step).astype(np.uint8))
    for dt in (np.int16, np.int32, np.int64):  # slow integer ramps
        out.append((1000 + np.arange(128 // np.dtype(dt).itemsize)).astype(dt).view(np.uint8))
    out.append(np.linspace(-1, 1, 32, dtype=np.float32).view(np.uint8))
    out.append(np.linspace(100, 101, 16, dtype=np.float64).view(np.uint8))
    for k in range(12):  # sparse bit patterns: few ones, to reach the cheap row patterns
        b = np.zeros(128, np.uint8)
        idx = rng.integers(0, 128, size=k + 1)
        b[idx] = (1 << rng.integers(0, 8, size=k + 1)).astype(np.uint8)
        out.append(b)
    return np.stack(out)


def golden_blocks():
    parts = [kat_blocks(), edge_blocks()]
    for kind in range(8):
        parts.append(synth(kind, 20221, 1000 * kind, 72, 1 << 20))
    blocks = np.concatenate(parts)
    assert blocks.shape[0] <= 640
    pad = synth("mixed_hashed", 5, 0, 640 - blocks.shape[0], 640)
    return np.concatenate([blocks, pad])


def gpgpusim_case():
    """Blocks and request types of the .log fixture (shared with tests/test_loader_gpgpu.py)."""
    blocks = golden_blocks()[:330]
    types = [(0, 4, 1, 0, 8, 4, 2, 0, 6, 5, 3, 0, 7)[i % 13] for i in range(len(blocks))]
    return blocks, types


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    blocks = golden_blocks()
    data = {"blocks": blocks}
    for cfg in CONFIGS:
        path = os.path.join(ROOT, "configs", cfg + ".json")
        ref = RefCompressor("VPC", path)
        sizes, sels = ref.compress(blocks)
        orig, comp, ratio = ref.totals()
        stat, fl, hist = ref.vpc_stats()
        data.update({f"{cfg}_sizes": sizes.astype(np.uint16), f"{cfg}_sels": sels.astype(np.int8),
                     f"{cfg}_totals": np.array([orig, comp], np.uint64), f"{cfg}_ratio": np.array([ratio]),
                     f"{cfg}_stat": stat, f"{cfg}_fl": fl,
                     f"{cfg}_hist_nz": np.argwhere(hist), f"{cfg}_hist_val": hist[hist != 0]})
        print(cfg, "ratio", ratio, "clusters", stat[:, 0].tolist())
    np.savez_compressed(os.path.join(gold, "mpc_blocks.npz"), **data)
    # CLI goldens: <tmp>/ds/golden_set.npy -> workload name "ds_golden_set" (main.cpp:142-157)
    with tempfile.TemporaryDirectory() as tmp:
        ds = os.path.join(tmp, "ds")
        os.makedirs(ds)
        np.save(os.path.join(ds, "golden_set.npy"), np.concatenate([blocks[:199], np.zeros((1, 128), np.uint8)]))
        for cfg in ("P6", "F4"):
            outdir = os.path.join(tmp, "out_" + cfg)
            os.makedirs(outdir)
            cfgpath = os.path.join(ROOT, "configs", cfg + ".json")
            stdout = ""
            for _ in range(2):
                r = subprocess.run([REF_BIN, "-a", "VPC", "-i", os.path.join(ds, "golden_set.npy"), "-c", cfgpath,
                                    "-o", outdir], capture_output=True, text=True, check=True)
                stdout = r.stdout
            for suffix in ("results.csv", "results_detail.csv"):
                with open(os.path.join(outdir, f"{cfg}_{suffix}")) as f, \
                        open(os.path.join(gold, f"cli_{cfg}_{suffix}"), "w") as g:
                    g.write(f.read())
            with open(os.path.join(gold, f"cli_{cfg}_stdout.txt"), "w") as g:
                g.write(stdout)
            print(cfg, "cli:", stdout.strip())
        # GPGPU-Sim trace (.log) through the reference CLI: 330 records of mixed request types (only GLOBAL_ACC_R/W are
        # compressed, main.cpp:222-224), the last record cut short (dropped: LoaderGPGPU.cpp:46-52)
        import hashlib
        import json
        log_blocks, log_types = gpgpusim_case()
        write_gpgpusim_log(os.path.join(ds, "trace_set.log"), log_blocks, log_types, truncate_last=5)
        outdir = os.path.join(tmp, "out_log")
        os.makedirs(outdir)
        r = subprocess.run([REF_BIN, "-a", "VPC", "-i", os.path.join(ds, "trace_set.log"), "-c", os.path.join(ROOT, "configs", "P6.json"),
                            "-o", outdir], capture_output=True, text=True, check=True)
        with open(os.path.join(outdir, "P6_results.csv")) as f, open(os.path.join(gold, "cli_log_P6_results.csv"), "w") as g:
            g.write(f.read())
        with open(os.path.join(outdir, "P6_results_detail.csv")) as f, open(os.path.join(gold, "cli_log_P6_results_detail.csv"), "w") as g:
            g.write(f.read())
        view = subprocess.run([REF_BIN, "-a", "VIEWER", "-i", os.path.join(ds, "trace_set.log")], capture_output=True, text=True, check=True).stdout
        view_npy = subprocess.run([REF_BIN, "-a", "VIEWER", "-i", os.path.join(ds, "golden_set.npy")], capture_output=True, text=True, check=True).stdout
        sc2 = subprocess.run([REF_BIN, "-a", "BDI", "-i", os.path.join(ds, "trace_set.log"), "-o", outdir], capture_output=True, text=True, check=True)
        with open(os.path.join(outdir, "BDI_results.csv")) as f:
            bdi_csv = f.read()
        with open(os.path.join(gold, "cli_log.json"), "w") as g:
            json.dump({"vpc_stdout": r.stdout, "viewer_sha256": hashlib.sha256(view.encode()).hexdigest(), "viewer_lines": view.count("\n"),
                       "viewer_head": view.splitlines()[:3], "viewer_npy_sha256": hashlib.sha256(view_npy.encode()).hexdigest(),
                       "viewer_npy_lines": view_npy.count("\n"), "bdi_stdout": sc2.stdout, "bdi_csv": bdi_csv}, g, indent=1)
        print("log cli:", r.stdout.strip(), "viewer lines", view.count("\n"), "bdi", sc2.stdout.strip())
        # secondary variants (config #5) through the reference CLI, on a 12 001-row dump so that SC2 leaves its
        # 10 000-line sampling phase (main.cpp:108-114); the dump is synth("mixed_hashed", 555, 0, 12000, 12000) + 1 row
        np.save(os.path.join(ds, "variants_set.npy"),
                np.concatenate([synth("mixed_hashed", 555, 0, 12000, 12000), np.zeros((1, 128), np.uint8)]))
        for alg in ("BDI", "FPC", "BPC", "CPACK", "SC2"):
            outdir = os.path.join(tmp, "out_" + alg)
            os.makedirs(outdir)
            r = subprocess.run([REF_BIN, "-a", alg, "-i", os.path.join(ds, "variants_set.npy"), "-o", outdir],
                               capture_output=True, text=True, check=True)
            with open(os.path.join(outdir, f"{alg}_results.csv")) as f, open(os.path.join(gold, f"cli_{alg}_results.csv"), "w") as g:
                g.write(f.read())
            with open(os.path.join(gold, f"cli_{alg}_stdout.txt"), "w") as g:
                g.write(r.stdout)
            print(alg, "cli:", r.stdout.strip())


if __name__ == "__main__":
    main()
