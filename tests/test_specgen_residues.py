"""The residue arithmetic the config compiler generates (csrc/mpc_specgen.cpp: byte gathers, shared subtracts, shift masks, the
16-bit-lane form of shifting predictors), compiled for the HOST and checked word by word against the per-byte definition of
PredictorModule.cpp:37-173 / ResidueModule.cpp:12-41 -- before any of it runs on a GPU.  tools/specgen --residues prints the
statements, tests/cpp/residue_probe_main.cpp holds the reference and the inputs (random bytes, borrow and sign-bit patterns)."""
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, random_config

SPECGEN = os.path.join(ROOT, "tools", "specgen")
CSRC = os.path.join(ROOT, "cal_22-mpc_b200", "csrc")
MAIN = os.path.join(ROOT, "tests", "cpp", "residue_probe_main.cpp")


def run_probe(cfg_path, tmp_path, env_extra=None):
    if not os.path.exists(SPECGEN):
        subprocess.run(["make", "-s", "tools/specgen"], cwd=ROOT, check=True)
    env = {k: v for k, v in os.environ.items() if not k.startswith("MPC_SPEC_")}
    env.update(env_extra or {})
    src = subprocess.run([SPECGEN, "--residues", str(cfg_path)], check=True, env=env, capture_output=True, text=True).stdout
    inc = tmp_path / "probe.inc"
    inc.write_text(src)
    exe = tmp_path / "probe"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", CSRC, f'-DPROBE_FILE="{inc}"', MAIN, "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    checked, bad = (int(v) for v in r.stdout.split())
    assert bad == 0 and checked > 0
    return src


@pytest.mark.parametrize("cfg", ["P6", "F4", "Z1", "E5", "S32", "S64"])
def test_shipped_configs(cfg, tmp_path):
    src = run_probe(os.path.join(ROOT, "configs", cfg + ".json"), tmp_path)
    if cfg == "P6":  # the Weight module takes the 16-bit-lane form: shifts and byte positions as multiply-adds
        assert "mpcdev::lanes_merge(" in src and "shiftmix(" not in src


def test_lane_form_can_be_switched_off(tmp_path):
    src = run_probe(os.path.join(ROOT, "configs", "P6.json"), tmp_path, {"MPC_SPEC_WLANES": "0"})
    assert "mpcdev::lanes_merge(" not in src and "shiftmix(" in src


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("L", [32, 64, 128])
def test_random_configs(seed, L, tmp_path):
    rng = np.random.default_rng(1000 * L + seed)
    cfg = random_config(rng, L=L, n_pred=4, table="pm" if seed % 2 else "cm")
    p = tmp_path / "cfg.json"
    p.write_text(json.dumps(cfg))
    run_probe(p, tmp_path)


@pytest.mark.parametrize("seed", range(10))
def test_structured_weight_predictors(seed, tmp_path):
    """Weight predictors as configs use them: a fixed stride back, shift distances periodic in the byte position -- every
    combination of source byte, target byte and shift the lane form distinguishes (top byte of a word, left shifts into the
    upper bytes, merged masks, net right shifts)."""
    rng = np.random.default_rng(77 + seed)
    L = int(rng.choice([32, 64, 128]))
    weights = [0.01, 0.1, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0, 100.0, 300.0]
    mods = [{"name": "AllZero"}, {"name": "AllWordSame"}]
    for _ in range(4):
        stride = int(rng.choice([1, 2, 3, 4, 5, 7, 8]))
        period = int(rng.choice([1, 2, 4, 8]))
        w = [float(rng.choice(weights)) for _ in range(period)]
        pred = {"name": "WeightBasePredictor", "LineSize": L, "RootIndex": 0,
                "BaseIndexTable": [0 if i < stride else i - stride for i in range(L)],
                "WeightTable": [w[i % period] for i in range(L)]}
        cols = list(range(1, L)) + [0]
        scan = {"TableSize": 8 * L, "Rows": [r for r in range(8) for _ in range(L)], "Cols": cols * 8}
        mods.append({"name": "PredComp", "submodules": {
            "ResidueModule": {"PredictorModule": pred}, "XORModule": {"consecutiveXOR": bool(rng.integers(0, 2))},
            "ScanModule": scan, "FPCModule": {"num_modules": 1, "0": {"name": "UncompressedPattern", "encodingBits": 17}}}})
    cfg = {"overview": {"num_modules": len(mods), "lineSize": L}, "modules": {str(i): m for i, m in enumerate(mods)}}
    p = tmp_path / "cfg.json"
    p.write_text(json.dumps(cfg))
    src = run_probe(p, tmp_path)
    assert "mpcdev::lanes_merge(" in src
