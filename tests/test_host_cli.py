"""The drop-in CLI (bin/compressor + bin/run): CSV bytes and stdout must equal what the unmodified reference
binary wrote for the same input (tests/golden/cli_*, produced by tests/golden/make_golden.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, cfg_path

BIN = os.path.join(ROOT, "bin", "compressor")
RUN = os.path.join(ROOT, "bin", "run")
GOLD = os.path.join(ROOT, "tests", "golden")


def _build():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", ROOT, "compressor"], check=True, capture_output=True)


def test_help_exits_zero_without_gpu():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True)  # no -i: help, exit 0 (main.cpp:50-71)
    assert r.returncode == 0 and "Usage" in r.stdout
    r = subprocess.run([BIN, "-a", "VPC", "-i", "x.npy"], capture_output=True, text=True)  # VPC without -c
    assert r.returncode == 0 and "Usage" in r.stdout


def test_bad_config_message_and_exit_code(tmp_path):
    _build()
    np.save(tmp_path / "d.npy", np.zeros((4, 128), np.uint8))
    r = subprocess.run([BIN, "-a", "VPC", "-i", str(tmp_path / "d.npy"), "-c", "/nonexistent.json", "-o", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 1 and 'Invalid File! "/nonexistent.json" is not valid path.' in r.stdout  # VPC.cpp:79


def test_format_double_matches_fmt():
    _build()
    import importlib
    # formatDouble lives in the host code; build a tiny shared object around it
    so = os.path.join(ROOT, "cal_22-mpc_b200", "host", "libmpcb_utils_test.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", os.path.join(ROOT, "cal_22-mpc_b200", "host", "utils.cpp"),
                    "-o", so], check=True)
    l = ctypes.CDLL(so)
    l.mpcb_format_double.argtypes = [ctypes.c_double, ctypes.c_char_p, ctypes.c_int]
    buf = ctypes.create_string_buffer(64)

    def f(v):
        assert l.mpcb_format_double(v, buf, 64) > 0
        return buf.value.decode()

    # values measured from fmt 12.1 in SURVEY.md section 7
    assert [f(v) for v in (0.0, 1.0, 1e15, 1e16, 1e-5, 1e-4, float("inf"))] == ["0", "1", "1000000000000000", "1e+16", "1e-05", "0.0001", "inf"]
    assert f(7.54837753741295) == "7.54837753741295" and f(341.3333333333333) == "341.3333333333333"
    assert f(1.4392912514797536) == "1.4392912514797536" and f(0.997078870496592) == "0.997078870496592"
    rng = np.random.default_rng(0)
    for v in np.concatenate([rng.random(2000) * 10.0 ** rng.integers(-8, 20, 2000), rng.integers(0, 10**9, 200).astype(float)]):
        v = float(v)
        r = repr(v)
        if "e" in r:
            mant, ex = r.split("e")
            want = mant.rstrip("0").rstrip(".") if "." in mant else mant
            want = f"{want}e{'-' if int(ex) < 0 else '+'}{abs(int(ex)):02d}"
        else:
            want = r[:-2] if r.endswith(".0") else r
        if 1e-4 <= abs(v) < 1e16:
            assert f(v) == want, (v, f(v), want)
        else:
            assert float(f(v)) == v


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["P6", "F4"])
def test_csv_bytes_equal_reference(tmp_path, golden, cfg):
    _build()
    ds = tmp_path / "ds"
    out = tmp_path / "out"
    ds.mkdir()
    out.mkdir()
    np.save(ds / "golden_set.npy", np.concatenate([golden["blocks"][:199], np.zeros((1, 128), np.uint8)]))
    stdout = ""
    for _ in range(2):  # appended twice, like the golden: header + 2 rows
        r = subprocess.run([BIN, "-a", "VPC", "-i", str(ds / "golden_set.npy"), "-c", cfg_path(cfg), "-o", str(out)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        stdout = r.stdout
    assert stdout == open(os.path.join(GOLD, f"cli_{cfg}_stdout.txt")).read()
    for suffix in ("results.csv", "results_detail.csv"):
        got = open(out / f"{cfg}_{suffix}").read()
        want = open(os.path.join(GOLD, f"cli_{cfg}_{suffix}")).read()
        assert got == want, suffix


@pytest.mark.gpu
def test_bin_run_walks_the_dataset_like_the_reference(tmp_path, golden):
    _build()
    ds = tmp_path / "data"
    out = tmp_path / "res"
    (ds / "splitA").mkdir(parents=True)
    (out / "splitA").mkdir(parents=True)
    blocks = golden["blocks"]
    np.save(ds / "splitA" / "train_set.npy", blocks[:101])
    np.save(ds / "top.npy", blocks[100:301])
    r = subprocess.run([RUN, "VPC", str(ds), str(out), cfg_path("P6")], capture_output=True, text=True,
                       cwd=os.path.join(ROOT, "bin"))
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "splitA : train_set" and lines[1].startswith("comp.ratio: ")
    assert lines[2] == "top.npy : top" or lines[2].endswith(": top")
    rows = open(out / "splitA" / "P6_results.csv").read().splitlines()
    assert rows[2].startswith("splitA_train_set,102400,")  # 100 blocks: the last row of the file is dropped
    assert open(out / "P6_results.csv").read().splitlines()[2].startswith("data_top,204800,")
    # single-line reference path still works: CompressLine through the same library
    want = golden["P6_sizes"][:100].astype(np.uint64).sum()
    assert int(rows[2].split(",")[2]) == int(want)


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ["BDI", "FPC", "BPC", "CPACK", "SC2"])
def test_variant_csv_bytes_equal_reference(tmp_path, alg):
    """BASELINE config #5 through the drop-in CLI: same CSV row and stdout as the reference binary on the same dump
    (tests/golden/cli_<ALG>_*, generated by tests/golden/make_golden.py from the unmodified reference)."""
    _build()
    from tools.gen_dump import synth
    ds = tmp_path / "ds"
    out = tmp_path / "out"
    ds.mkdir()
    out.mkdir()
    np.save(ds / "variants_set.npy", np.concatenate([synth("mixed_hashed", 555, 0, 12000, 12000), np.zeros((1, 128), np.uint8)]))
    r = subprocess.run([BIN, "-a", alg, "-i", str(ds / "variants_set.npy"), "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout == open(os.path.join(GOLD, f"cli_{alg}_stdout.txt")).read()
    assert open(out / f"{alg}_results.csv").read() == open(os.path.join(GOLD, f"cli_{alg}_results.csv")).read()


@pytest.mark.gpu
def test_multi_gpu_cli_equals_reference(tmp_path, golden):
    """--gpus 2: contiguous shards on two devices, statistics all-reduced with NCCL; the CSV bytes must
    still be the reference's.  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _build()
    ds = tmp_path / "ds"
    out = tmp_path / "out"
    ds.mkdir()
    out.mkdir()
    np.save(ds / "golden_set.npy", np.concatenate([golden["blocks"][:199], np.zeros((1, 128), np.uint8)]))
    stdout = ""
    for _ in range(2):
        r = subprocess.run([BIN, "-a", "VPC", "-i", str(ds / "golden_set.npy"), "-c", cfg_path("F4"), "-o", str(out), "--gpus", "2"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        stdout = r.stdout
    assert stdout == open(os.path.join(GOLD, "cli_F4_stdout.txt")).read()
    for suffix in ("results.csv", "results_detail.csv"):
        assert open(out / f"F4_{suffix}").read() == open(os.path.join(GOLD, f"cli_F4_{suffix}")).read(), suffix


def test_reference_style_plugins_compile_and_run_unchanged(tmp_path):
    """Compressor.h / Loader.h boundary: a compressor that only implements CompressLine and a loader that only implements
    the four reference virtuals compile against the mirror headers; the default CompressBatch / GetChunk bodies are the
    reference's per-line loops (main.cpp:229-244), including the dropped isEnd line."""
    host = os.path.join(ROOT, "cal_22-mpc_b200", "host")
    exe = str(tmp_path / "refstyle")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + host, "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "ref_style_plugins.cpp"), os.path.join(host, "utils.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, check=True)
    assert r.stdout.split() == ["9", str(9 * 32 * 8), str(sum(range(9)) * 32)]


@pytest.mark.gpu
def test_line_size_mismatch_is_refused(tmp_path):
    """A [N,32] dump under a 128-byte config would be read out of bounds by the reference (VPC.cpp:101 ignores the loader's
    line size): refused with a message and exit code 1."""
    _build()
    np.save(tmp_path / "d32.npy", np.zeros((64, 32), np.uint8))
    r = subprocess.run([BIN, "-a", "VPC", "-i", str(tmp_path / "d32.npy"), "-c", cfg_path("P6"), "-o", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Line size mismatch" in r.stdout


@pytest.mark.gpu
def test_config1_survey_dump_through_bin_run(tmp_path):
    """BASELINE configs[0]: `bin/run VPC <ds> <out> P6.json` over the survey's 64 MiB mixed dump (524 289 rows; BASELINE.md
    section 3) must print the reference's `comp.ratio: 1.4392912514797536` and write the CSV bytes the unmodified reference
    wrote for the same dump through its own bin/run (tests/golden/cli_survey_*, made by tests/golden/make_golden_survey.py)."""
    import hashlib
    import json
    from tools.gen_dump import survey_mixed
    _build()
    meta = json.load(open(os.path.join(GOLD, "cli_survey.json")))
    dump = survey_mixed(meta["rows"])
    # the recipe draws from numpy's default_rng: the golden only applies to the same byte stream
    assert hashlib.sha256(dump.tobytes()).hexdigest() == meta["sha256"], "numpy's random stream differs from the golden's"
    ds, out = tmp_path / "ds", tmp_path / "out"
    ds.mkdir()
    out.mkdir()
    np.save(ds / "survey_mixed.npy", dump)
    r = subprocess.run([RUN, "VPC", str(ds), str(out), cfg_path("P6")], capture_output=True, text=True, cwd=os.path.join(ROOT, "bin"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "comp.ratio: 1.4392912514797536" in r.stdout
    assert r.stdout == open(os.path.join(GOLD, "cli_survey_P6_stdout.txt")).read()
    for suffix in ("results.csv", "results_detail.csv"):
        assert open(out / f"P6_{suffix}").read() == open(os.path.join(GOLD, f"cli_survey_P6_{suffix}")).read(), suffix
