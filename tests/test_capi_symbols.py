"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/mpc_capi.h declares, the
config front end agrees with an independent parse of the same JSON, and errors are reported, not fatal."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from helpers import ROOT, SHIPPED, cfg_path, random_config


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpc_capi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(mpcb):
    names = declared_symbols()
    assert len(names) >= 19
    l = ctypes.CDLL(mpcb.LIB_PATH)
    for n in names:
        assert hasattr(l, n), n
    assert sorted(mpcb.SYMBOLS) == names
    assert b"sm_100a" in mpcb.lib().mpc_version()


def test_layout_constants_agree(mpcb):
    text = open(os.path.join(ROOT, "include", "mpc_capi.h")).read()
    assert int(re.search(r"#define MPC_MAX_LINE (\d+)", text).group(1)) == mpcb.capi.MAX_LINE
    assert int(re.search(r"#define MPC_MAX_MODULES (\d+)", text).group(1)) == mpcb.capi.MAX_MODULES
    # the library validates a zeroed pod of the size Python believes in without touching memory past it
    pod = mpcb.ConfigPod()
    err = ctypes.create_string_buffer(256)
    assert mpcb.lib().mpc_config_validate(ctypes.byref(pod), err, 256) == -2


@pytest.mark.parametrize("cfg", SHIPPED)
def test_parser_matches_python_json(mpcb, cfg):
    pod = mpcb.load_config(path=cfg_path(cfg))
    ref = json.load(open(cfg_path(cfg)))
    check_pod(pod, ref)


@pytest.mark.parametrize("seed", range(8))
def test_parser_random_configs(mpcb, seed):
    rng = np.random.default_rng(seed)
    ref = random_config(rng, L=int(rng.choice([32, 64, 128])))
    check_pod(mpcb.load_config(text=json.dumps(ref, indent=1)), ref)


def check_pod(pod, ref):
    ov = ref["overview"]
    n, L = ov["num_modules"], ov["lineSize"]
    assert (pod.num_modules, pod.line_size) == (n, L)
    names = [ref["modules"][str(i)]["name"] for i in range(n)]
    has_ws = int(any(x in ("AllWordSame", "ByteplaneAllSame") for x in names))
    assert pod.has_wordsame == has_ws and pod.first_predcomp == 1 + has_ws
    if "encoding_bits" in ov:
        assert list(pod.enc_bits[: n + 1]) == ov["encoding_bits"][: n + 1]
    else:
        assert list(pod.enc_bits[: n + 1]) == [int(np.ceil(np.log2(np.float32(n + 1))))] * (n + 1)
    for i in range(pod.first_predcomp, n):
        m, sub = pod.modules[i], ref["modules"][str(i)]["submodules"]
        ps, sc = sub["ResidueModule"]["PredictorModule"], sub["ScanModule"]
        assert m.kind == 2 and m.root == ps["RootIndex"]
        assert m.consecutive_xor == int(sub["XORModule"]["consecutiveXOR"])
        assert m.table_size == sc["TableSize"]
        assert list(m.scan_row[: m.table_size]) == sc["Rows"] and list(m.scan_col[: m.table_size]) == sc["Cols"]
        if ps["name"] == "DiffBasePredictor":
            assert list(m.base[:L]) == ps["BaseIndexTable"]
            assert list(m.diff[:L]) == [d & 0xFF for d in ps["DiffTable"]]
        if ps["name"] == "WeightBasePredictor":
            want = [max(-8, min(8, int(np.log2(np.float32(w))))) for w in ps["WeightTable"]]
            got = list(m.shift[:L])
            for j in range(L):
                if j != m.root:
                    assert got[j] == want[j], (j, ps["WeightTable"][j])


BAD = [
    ({"overview": {"num_modules": 1, "lineSize": 128}, "modules": {"0": {"name": "Nope"}}}, "is not a valid compression module"),
    ({"overview": {"num_modules": 1, "lineSize": 128}, "modules": {"0": {"name": "PredComp", "submodules": {
        "ResidueModule": {"PredictorModule": {"name": "Magic", "LineSize": 128, "RootIndex": 0}}}}}}, "is not a valid predictor module"),
    ({"overview": {"num_modules": 1, "lineSize": 100}, "modules": {"0": {"name": "AllZero"}}}, "lineSize"),
    ({"overview": {"num_modules": 0, "lineSize": 128}, "modules": {}}, "num_modules"),
    ({"overview": {"num_modules": 2, "lineSize": 128}, "modules": {"0": {"name": "AllWordSame"}, "1": {"name": "AllZero"}}}, "module 0 must be AllZero"),
]


@pytest.mark.parametrize("cfg,msg", BAD)
def test_bad_configs_are_rejected_with_a_message(mpcb, cfg, msg):
    with pytest.raises(mpcb.MpcError) as e:
        mpcb.load_config(text=json.dumps(cfg))
    assert msg in str(e.value)


def test_syntax_error_and_missing_file(mpcb):
    with pytest.raises(mpcb.MpcError) as e:
        mpcb.load_config(text="{ not json")
    assert "JSON syntax error" in str(e.value)
    with pytest.raises(mpcb.MpcError) as e:
        mpcb.load_config(path="/nonexistent/cfg.json")
    assert "is not valid path" in str(e.value)  # VPC.cpp:79


def test_create_without_gpu_fails_loudly(mpcb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mpcb.MpcError) as e:
        mpcb.Mpc(cfg_path("P6"))
    assert "no CPU path" in str(e.value)


def test_stats_expand_host_side(mpcb):
    # mpc_stats_expand is pure host code: feed it a hand-built statistics vector
    pod = mpcb.load_config(path=cfg_path("P6"))
    K, HB = mpcb.capi.MAX_MODULES + 1, mpcb.capi.HIST_BINS
    w = np.zeros(mpcb.capi.STATS_WORDS, dtype=np.uint64)
    w[2 * K + 1 * HB + 3] = 5      # cluster 0 (AllZero): five blocks of 3 bits
    w[2 * K + 0 * HB + 1027] = 2   # uncompressed: two blocks of 1027 bits
    w[0], w[K] = 1000, 70000
    out = mpcb.StatsPod()
    assert mpcb.lib().mpc_stats_expand(ctypes.byref(pod), w.ctypes.data, w.size, ctypes.byref(out)) == 0
    st = mpcb.Stats(out, pod.num_modules, pod.line_size)
    assert st.blocks == 7 and st.OriginalSize == 7 * 1024 and st.CompressedSize == 15 + 2054
    assert st.count[:2].tolist() == [2, 5] and st.res_lines[:2].tolist() == [2, 0]
    assert st.mae(0) == (1000 / 128) / 2
