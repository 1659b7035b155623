"""The config compiler's structural decisions (csrc/mpc_specgen.cpp), checked on the generated sources without a GPU:
which form each shipped config gets, and that the committed ahead-of-time sources are what the generator emits today."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, random_config

SPECGEN = os.path.join(ROOT, "tools", "specgen")


@pytest.fixture(scope="module")
def generated(tmp_path_factory):
    if not os.path.exists(SPECGEN):
        subprocess.run(["make", "-s", "tools/specgen"], cwd=ROOT, check=True)
    out = tmp_path_factory.mktemp("spec")
    cfgs = [os.path.join(ROOT, "configs", n + ".json") for n in ("P6", "F4", "Z1", "E5", "S32", "S64")]
    rng = np.random.default_rng(3)
    perm = out / "PERM.json"
    perm.write_text(json.dumps(random_config(rng, L=64, n_pred=2, table="perm")))
    env = {k: v for k, v in os.environ.items() if not k.startswith("MPC_SPEC_")}
    subprocess.run([SPECGEN, str(out)] + cfgs + [str(perm)], check=True, env=env, capture_output=True)
    return {n: (out / f"spec_{n}.cu").read_text() for n in ("P6", "F4", "Z1", "E5", "S32", "S64", "PERM")}


def test_committed_sources_are_current(generated):
    for n in ("P6", "F4", "Z1", "E5", "S32", "S64"):
        committed = open(os.path.join(ROOT, "cal_22-mpc_b200", "csrc", "spec", f"spec_{n}.cu")).read()
        assert committed == generated[n], f"csrc/spec/spec_{n}.cu is stale: run `make` (tools/specgen) and commit"


def test_plane_major_probe_config_takes_the_state_machine_form(generated):
    src = generated["P6"]
    assert "kPm2 = true" in src and "select_encode(uint32_t (&x)[32]" in src
    for m in (2, 3, 4, 5):  # one residue pass per module, no per-module winner pass
        assert f"uint32_t res_{m}(" in src and f"full_{m}(" not in src
    assert src.count("pm2_tail(") == 2 and "encode_pm<8," in src  # defined once, called once


def test_plane_major_probe_config_trades_alu_for_fma_instructions(generated):
    """P6's Weight module takes the 16-bit-lane residue (shifts and byte positions as multiply-adds), its canonical layout the 4 x 4
    byte transposes; the plain-copy predictors keep the byte-wise subtract (the lane form measured slower there), and E5 -- one
    plane-major module among column-major ones, at its register limit -- keeps the per-word gathers."""
    src = generated["P6"]
    res5 = src[src.index("uint32_t res_5("):src.index("// leading zero rows from complete residues")]
    assert res5.count("mpcdev::lanes_merge(") == 32 and "shiftmix(" not in res5 and "sub_u8x4(" not in res5
    res2 = src[src.index("uint32_t res_2("):src.index("uint32_t res_3(")]
    assert "lanes_merge" not in res2 and res2.count("sub_u8x4(") == 32
    tail = src[src.index("uint32_t pm2_tail("):src.index("struct Cfg {")]
    assert tail.count("0x5140u") == 16 and tail.count("0x7362u") == 16  # eight groups of four residue words, both stages
    assert "0x5140u" not in generated["E5"]


def test_column_major_config_gets_lut_paired_rows_and_selector_groups(generated):
    src = generated["F4"]
    assert "kUseLut = true" in src and "kLutXor = 1" in src and "kWarps = 20" in src
    assert "row_pair_step(" in src  # paired row layout
    assert "selector group 0: modules 2, 5" in src and "selector group 1: modules 3, 4" in src
    assert "full_g0(x, c, sa, sq, s0, s1)" in src
    assert "kAdaptiveEncode = true" in src


def test_short_lines_and_arbitrary_tables(generated):
    assert "kLineBytes = 32" in generated["S32"] and "kWords = 8" in generated["S32"]
    assert "kLineBytes = 64" in generated["S64"] and "kWords = 16" in generated["S64"]
    perm = generated["PERM"]
    assert "bit-gather (arbitrary table)" in perm and "kLineBytes = 64" in perm and "kLutXor = 0" in perm
    assert "kPm2 = false" in generated["E5"]  # mixed families keep the per-module passes
