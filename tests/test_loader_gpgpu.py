"""GPGPU-Sim trace loader (.log, reference src/loader/LoaderGPGPU.cpp namespace gpgpusim) behind the drop-in CLI.
Goldens (tests/golden/cli_log*) were written by the unmodified reference binary on the same trace
(tests/golden/make_golden.py: 330 records of mixed request types, last record cut short)."""
import hashlib
import importlib.util
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, cfg_path
from tools.gen_dump import write_gpgpusim_log

BIN = os.path.join(ROOT, "bin", "compressor")
GOLD = os.path.join(ROOT, "tests", "golden")


def _build():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", ROOT, "compressor"], check=True, capture_output=True)


def _case():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.gpgpusim_case()


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLD, "cli_log.json")))


def test_viewer_on_trace_equals_reference(tmp_path, gold):
    """VIEWER is host-only (main.cpp:121-125, 250-300): record walk, request-type filter, truncated last record."""
    _build()
    blocks, types = _case()
    ds = tmp_path / "ds"
    ds.mkdir()
    write_gpgpusim_log(ds / "trace_set.log", blocks, types, truncate_last=5)
    r = subprocess.run([BIN, "-a", "VIEWER", "-i", str(ds / "trace_set.log")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("\n") == gold["viewer_lines"]
    assert r.stdout.splitlines()[:3] == gold["viewer_head"]
    assert hashlib.sha256(r.stdout.encode()).hexdigest() == gold["viewer_sha256"]
    # a trace that ends on a record boundary keeps its last record
    write_gpgpusim_log(ds / "whole.log", blocks[:10], [0] * 10)
    r = subprocess.run([BIN, "-a", "VIEWER", "-i", str(ds / "whole.log")], capture_output=True, text=True)
    assert r.stdout.count("\n") == 10 and r.stdout.startswith("R: ")
    # header only: no lines
    write_gpgpusim_log(ds / "empty.log", blocks[:0], [])
    r = subprocess.run([BIN, "-a", "VIEWER", "-i", str(ds / "empty.log")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == ""


def test_viewer_on_npy_equals_reference(tmp_path, golden, gold):
    _build()
    ds = tmp_path / "ds"
    ds.mkdir()
    np.save(ds / "golden_set.npy", np.concatenate([golden["blocks"][:199], np.zeros((1, 128), np.uint8)]))
    r = subprocess.run([BIN, "-a", "VIEWER", "-i", str(ds / "golden_set.npy")], capture_output=True, text=True)
    assert r.returncode == 0
    assert r.stdout.count("\n") == gold["viewer_npy_lines"]  # the last row is never shown (LoaderNPY.cpp:28-32)
    assert hashlib.sha256(r.stdout.encode()).hexdigest() == gold["viewer_npy_sha256"]


def test_trace_header_errors(tmp_path):
    _build()
    bad = tmp_path / "bad.log"
    bad.write_bytes(bytes([16]) + b"\0" * 200)  # key count != 17 (LoaderGPGPU.cpp:91-95)
    r = subprocess.run([BIN, "-a", "VIEWER", "-i", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "The header of the GPGPU-sim trace file is not valid." in r.stdout
    short = tmp_path / "short.log"
    short.write_bytes(bytes([17]) + b"\0" * 50)  # header runs into the end of the file (LoaderGPGPU.cpp:108-112)
    r = subprocess.run([BIN, "-a", "VIEWER", "-i", str(short)], capture_output=True, text=True)
    assert r.returncode == 1 and "not valid" in r.stdout
    r = subprocess.run([BIN, "-a", "VIEWER", "-i", str(tmp_path / "missing.log")], capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to open a file." in r.stdout  # LoaderGPGPU.cpp:84-88


@pytest.mark.gpu
def test_trace_csv_bytes_equal_reference(tmp_path, gold):
    _build()
    blocks, types = _case()
    ds = tmp_path / "ds"
    out = tmp_path / "out"
    ds.mkdir()
    out.mkdir()
    write_gpgpusim_log(ds / "trace_set.log", blocks, types, truncate_last=5)
    r = subprocess.run([BIN, "-a", "VPC", "-i", str(ds / "trace_set.log"), "-c", cfg_path("P6"), "-o", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout == gold["vpc_stdout"]
    for suffix in ("results.csv", "results_detail.csv"):
        assert open(out / f"P6_{suffix}").read() == open(os.path.join(GOLD, f"cli_log_P6_{suffix}")).read(), suffix
    r = subprocess.run([BIN, "-a", "BDI", "-i", str(ds / "trace_set.log"), "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout == gold["bdi_stdout"] and open(out / "BDI_results.csv").read() == gold["bdi_csv"]


@pytest.mark.gpu
def test_mixed_size_trace_is_refused(tmp_path):
    _build()
    blocks, _ = _case()
    import struct
    p = tmp_path / "mixed.log"
    write_gpgpusim_log(p, blocks[:4], [0, 0, 0, 0])
    with open(p, "ab") as f:  # one more GLOBAL_ACC_R record with a 32-byte payload
        f.write(struct.pack("<BBQIIIIIQIIIIII", 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 32) + bytes(32))
    r = subprocess.run([BIN, "-a", "VPC", "-i", str(p), "-c", cfg_path("P6"), "-o", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "mixed-size traces are not supported" in r.stdout


@pytest.mark.gpu
def test_sector_trace_every_algorithm_equals_the_reference_binary(tmp_path):
    """GPGPU-Sim traces are 32-byte sectors (ACCESS_GRAN 32): `compressor -a <alg> -i trace.log` for every GPU algorithm on such a trace,
    stdout and CSV bytes against the unmodified reference binary run on the same file (oracle/_ref/compressor travels prebuilt)."""
    from oracle.bridge import REF_BIN
    from tools.gen_dump import synth
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/compressor not built")
    _build()
    sectors = synth("mixed_hashed", 41, 0, 700, 700).reshape(-1, 32)[:2777]
    types = [(0, 4, 1, 0, 2, 4)[i % 6] for i in range(sectors.shape[0])]  # GLOBAL_ACC_R / W kept, others skipped (main.cpp:222-224)
    ds = tmp_path / "ds"
    ds.mkdir()
    write_gpgpusim_log(ds / "sectors_set.log", sectors, types, truncate_last=3)
    # (FPC is left out: the reference's zero-run scan reads past the line, FPC.cpp:26 -- on short lines its statistics are inflated
    #  non-deterministically and the binary can crash; FPC's 32-byte path is pinned through the oracle in tests/test_variants.py)
    for alg, cfg in (("VPC", cfg_path("S32")), ("BDI", None), ("BPC", None), ("SC2", None), ("PATTERN", None)):
        outs = []
        for exe, name in ((BIN, "ours"), (REF_BIN, "ref")):
            out = tmp_path / f"{alg}_{name}"
            out.mkdir()
            cmd = [exe, "-a", alg, "-i", str(ds / "sectors_set.log"), "-o", str(out)] + (["-c", cfg] if cfg else [])
            r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(out))
            assert r.returncode == 0, (alg, name, r.stdout + r.stderr)
            files = {f: open(out / f).read() for f in sorted(os.listdir(out))}
            outs.append((r.stdout, files))
        assert outs[0][0] == outs[1][0], (alg, outs[0][0], outs[1][0])
        assert outs[0][1].keys() == outs[1][1].keys() and len(outs[0][1]) >= 1, alg
        for f in outs[0][1]:
            assert outs[0][1][f] == outs[1][1][f], (alg, f)
