"""Pins the CPU oracle (oracle/mpc_oracle.c) to the reference: committed golden vectors generated from the
unmodified reference build (tests/golden/make_golden.py) and the known answers of SURVEY.md section 8c."""
import numpy as np
import pytest

from helpers import SHIPPED, cfg_path
from oracle.bridge import OracleMPC
from tools.gen_dump import kat_blocks, survey_mixed


def test_known_answer_blocks_p6():
    # SURVEY.md section 8c table, column "VPC size (sel)" under configs/P6.json
    r = OracleMPC(cfg_path("P6")).run(kat_blocks())
    assert r.sizes.tolist() == [3, 35, 146, 1027, 1027, 1027, 201, 263, 35]
    assert r.sels.tolist() == [0, 1, 4, -1, -1, -1, 5, 3, 1]
    # 9-row file total through the reference's main.cpp (row 8 dropped): 8192 / 3729
    assert int(r.sizes[:8].sum()) == 3729


@pytest.mark.parametrize("cfg", SHIPPED)
def test_oracle_matches_reference_golden(golden, cfg):
    o = OracleMPC(cfg_path(cfg))
    r = o.run(golden["blocks"])
    assert np.array_equal(r.sizes, golden[f"{cfg}_sizes"].astype(np.uint32))
    assert np.array_equal(r.sels, golden[f"{cfg}_sels"].astype(np.int32))
    orig, comp = (int(v) for v in golden[f"{cfg}_totals"])
    assert (r.OriginalSize, r.CompressedSize) == (orig, comp)
    stat, fl = golden[f"{cfg}_stat"], golden[f"{cfg}_fl"]
    assert np.array_equal(r.count, stat[:, 0])
    assert np.array_equal(r.comp_bits, stat[:, 2])
    assert np.array_equal(r.res_lines, stat[:, 3])
    L = o.L
    for k in range(o.n + 1):  # MAE / MSE doubles, bit for bit (SURVEY.md section 7 "MAE/MSE exactly")
        n = int(r.res_lines[k])
        mae = (float(int(r.res_abs[k])) / L) / float(n) if n else 0.0
        mse = (float(int(r.res_sq[k])) / L) / float(n) if n else 0.0
        assert mae == fl[k, 1] and mse == fl[k, 2]
    hist = np.zeros_like(r.hist)
    nzi = golden[f"{cfg}_hist_nz"]
    hist[nzi[:, 0], nzi[:, 1]] = golden[f"{cfg}_hist_val"]
    assert np.array_equal(r.hist, hist)


def test_survey_mixed_80000():
    # n = 80000 rows of the survey recipe under P6 -> sizes sum 56 932 369 bits
    d = survey_mixed(80001)[:80000]
    r = OracleMPC(cfg_path("P6")).run(d)
    assert r.CompressedSize == 56932369
    assert r.OriginalSize / r.CompressedSize == pytest.approx(1.43890025, abs=1e-8)
