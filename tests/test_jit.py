"""Run-time specialisation: the config compiler (csrc/mpc_specgen.cpp) + NVRTC build a thread-per-block kernel for any
config with 32-, 64- or 128-byte lines and column-/plane-major scan tables.  CPU: the generated sources compile for sm_100a.
GPU: the kernels are bit-exact against the oracle."""
import json

import numpy as np
import pytest

from helpers import SHIPPED, cfg_path, random_blocks, random_config
from oracle.bridge import OracleMPC


def eligible_config(seed, L=128):
    rng = np.random.default_rng(seed)
    mode = ["cm", "pm", "pmr", "cms", "cm", "pm"][seed % 6]
    cfg = random_config(rng, L=L, n_pred=int(rng.integers(1, 5)), table=mode)
    if seed % 4 == 3:  # mixed families in one config
        other = random_config(rng, L=L, n_pred=1, table="pm" if mode.startswith("cm") else "cm")
        first = [k for k in sorted(other["modules"], key=int) if other["modules"][k]["name"] == "PredComp"][0]
        n = cfg["overview"]["num_modules"]
        cfg["modules"][str(n)] = other["modules"][first]
        cfg["overview"]["num_modules"] = n + 1
        if "encoding_bits" in cfg["overview"]:
            cfg["overview"]["encoding_bits"].append(3)
    return cfg


@pytest.mark.parametrize("cfg", SHIPPED)
def test_shipped_configs_compile_with_nvrtc(mpcb, cfg):
    rc, nbytes, log = mpcb.jit_compile_check(cfg_path(cfg))
    assert rc == 0 and nbytes > 10000, log


@pytest.mark.parametrize("seed", range(6))
def test_random_eligible_configs_compile(mpcb, seed):
    pod = mpcb.load_config(text=json.dumps(eligible_config(seed)))
    rc, nbytes, log = mpcb.jit_compile_check(pod)
    assert rc == 0 and nbytes > 10000, log


def test_ineligible_configs_are_reported(mpcb, monkeypatch):
    rng = np.random.default_rng(1)
    monkeypatch.setenv("MPC_SPEC_BITGATHER", "0")  # arbitrary scan tables left to the generic kernel
    pod = mpcb.load_config(text=json.dumps(random_config(rng, L=128, n_pred=2, table="perm")))
    rc, _, log = mpcb.jit_compile_check(pod)
    assert rc == -2 and "neither column-major nor plane-major" in log


@pytest.mark.parametrize("L", [32, 128])
@pytest.mark.parametrize("table", ["perm", "dup", "short"])
def test_arbitrary_scan_tables_compile(mpcb, table, L):
    """any (row, column) table (ScanModule.cpp:6-22) -- permutations, duplicates, short tables: the bit-gather family"""
    rng = np.random.default_rng(L + len(table))
    pod = mpcb.load_config(text=json.dumps(random_config(rng, L=L, n_pred=2, table=table)))
    rc, nbytes, log = mpcb.jit_compile_check(pod)
    assert rc == 0 and nbytes > 5000, log


def shared_scan_pm_config(seed, L):
    """Every module plane-major with ONE column order, planes MSB first (the shape of the survey's probe config P6): the
    state-machine form of the plane-major path (select_encode)."""
    rng = np.random.default_rng(seed)
    cfg = random_config(rng, L=L, n_pred=int(rng.integers(1, 5)), table="pm")
    mods = [m for m in cfg["modules"].values() if m["name"] == "PredComp"]
    for m in mods[1:]:
        m["submodules"]["ScanModule"] = mods[0]["submodules"]["ScanModule"]
    return cfg


@pytest.mark.parametrize("L", [32, 64])
@pytest.mark.parametrize("seed", range(6))
def test_short_line_configs_compile(mpcb, seed, L):
    pod = mpcb.load_config(text=json.dumps(eligible_config(seed, L)))
    rc, nbytes, log = mpcb.jit_compile_check(pod)
    assert rc == 0 and nbytes > 5000, log


@pytest.mark.parametrize("L", [32, 64, 128])
@pytest.mark.parametrize("seed", range(3))
def test_shared_scan_plane_major_configs_compile(mpcb, seed, L):
    pod = mpcb.load_config(text=json.dumps(shared_scan_pm_config(seed, L)))
    rc, nbytes, log = mpcb.jit_compile_check(pod)
    assert rc == 0 and nbytes > 5000, log


def check_jit_config(mpcb, cfg, seed, L=128, n=3000):
    m = mpcb.Mpc(mpcb.load_config(text=json.dumps(cfg)))
    assert m.kernel_name() == "spec_thread:jit"
    rng = np.random.default_rng(seed)
    blocks = random_blocks(rng, n, L)
    sizes, sels, st = m.compress(blocks)
    r = OracleMPC(cfg).run(blocks)
    bad = np.nonzero((sizes != r.sizes) | (sels != r.sels))[0]
    assert bad.size == 0, (bad[:5], sizes[bad[:5]], r.sizes[bad[:5]], sels[bad[:5]], r.sels[bad[:5]])
    assert st.CompressedSize == r.CompressedSize and np.array_equal(st.count, r.count)
    assert np.array_equal(st.res_abs, r.res_abs) and np.array_equal(st.res_sq, r.res_sq)
    hb = r.hist.shape[1]
    assert np.array_equal(st.hist[:, :hb], r.hist)


@pytest.mark.gpu
@pytest.mark.parametrize("L", [32, 64])
@pytest.mark.parametrize("seed", range(12))
def test_short_line_jit_kernels_match_oracle(mpcb, seed, L):
    """32- and 64-byte lines on the thread-per-block kernel: four / two lines per thread; ragged count on purpose"""
    check_jit_config(mpcb, eligible_config(200 + seed, L), seed, L, n=3001 + seed)


@pytest.mark.gpu
@pytest.mark.parametrize("L", [32, 64, 128])
@pytest.mark.parametrize("table", ["perm", "dup", "short", "random"])
@pytest.mark.parametrize("seed", range(2))
def test_arbitrary_scan_table_jit_kernels_match_oracle(mpcb, seed, table, L):
    """arbitrary scan tables on the thread-per-block kernel (bit-gather family), also mixed with the other families"""
    rng = np.random.default_rng(1000 * L + 10 * seed + len(table))
    cfg = random_config(rng, L=L, n_pred=int(rng.integers(1, 4)), table=table)
    check_jit_config(mpcb, cfg, seed, L, n=2000 + seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_plane_major_deferral_buffer_variant_matches_oracle(mpcb, monkeypatch, seed):
    """MPC_SPEC_DEFER=1 (profiles/r02_defer.txt: measured, off by default): blocks whose winner is not the last module are parked in a
    per-warp buffer and finished in batches -- every block still exactly once, ragged tail included"""
    monkeypatch.setenv("MPC_SPEC_DEFER", "1")
    check_jit_config(mpcb, shared_scan_pm_config(400 + seed, 128), seed, 128, n=5003)


@pytest.mark.gpu
@pytest.mark.parametrize("L", [32, 64, 128])
@pytest.mark.parametrize("seed", range(4))
def test_shared_scan_plane_major_jit_kernels_match_oracle(mpcb, seed, L):
    check_jit_config(mpcb, shared_scan_pm_config(300 + seed, L), seed, L, n=2999)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_jit_kernels_match_oracle(mpcb, seed):
    cfg = eligible_config(100 + seed)
    m = mpcb.Mpc(mpcb.load_config(text=json.dumps(cfg)))
    assert m.kernel_name() == "spec_thread:jit"
    rng = np.random.default_rng(seed)
    blocks = random_blocks(rng, 3000)
    sizes, sels, st = m.compress(blocks)
    r = OracleMPC(cfg).run(blocks)
    bad = np.nonzero((sizes != r.sizes) | (sels != r.sels))[0]
    assert bad.size == 0, (bad[:5], sizes[bad[:5]], r.sizes[bad[:5]], sels[bad[:5]], r.sels[bad[:5]])
    assert st.CompressedSize == r.CompressedSize and np.array_equal(st.count, r.count)
    assert np.array_equal(st.res_abs, r.res_abs) and np.array_equal(st.res_sq, r.res_sq)
    m.set_kernel(1)  # the generic kernel agrees, too
    sizes2, sels2, _ = m.compress(blocks)
    assert np.array_equal(sizes2, sizes) and np.array_equal(sels2, sels)


@pytest.mark.gpu
def test_jit_can_be_disabled(mpcb, monkeypatch):
    monkeypatch.setenv("MPC_JIT", "0")
    m = mpcb.Mpc(mpcb.load_config(text=json.dumps(eligible_config(7))))
    assert m.kernel_name() == "generic_warp"
    with pytest.raises(mpcb.MpcError) as e:
        m.set_kernel(2)
    assert "MPC_JIT=0" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 4])
def test_jit_kernels_with_the_tma_tile_loader_match_oracle(mpcb, monkeypatch, seed):
    """MPC_SPEC_TMA=1: tiles arrive by cp.async.bulk.tensor (2D tensor map, 128-byte swizzle, per-warp mbarrier) instead
    of cp.async; ragged block counts exercise the zero-filled rows past the end of the dump."""
    monkeypatch.setenv("MPC_SPEC_TMA", "1")
    monkeypatch.setenv("MPC_JIT_CACHE_DIR", "")
    cfg = eligible_config(200 + seed)
    m = mpcb.Mpc(mpcb.load_config(text=json.dumps(cfg)))
    assert m.kernel_name() == "spec_thread:jit"
    rng = np.random.default_rng(seed)
    for n in (1, 31, 33, 2999):
        blocks = random_blocks(rng, n)
        sizes, sels, st = m.compress(blocks)
        r = OracleMPC(cfg).run(blocks)
        assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels), n
        assert st.CompressedSize == r.CompressedSize and np.array_equal(st.res_abs, r.res_abs)


def test_tma_variant_compiles_with_nvrtc(mpcb, monkeypatch):
    monkeypatch.setenv("MPC_SPEC_TMA", "1")
    rc, nbytes, log = mpcb.jit_compile_check(cfg_path("F4"))
    assert rc == 0 and nbytes > 10000, log


@pytest.mark.parametrize("flags", [{"MPC_SPEC_REGROUP": "1"}, {"MPC_SPEC_REGROUP": "1", "MPC_SPEC_FUSED": "0"}, {"MPC_SPEC_ADAPTIVE": "0"},
                                   {"MPC_SPEC_FUSED": "1"}, {"MPC_SPEC_DEFER": "1"}, {"MPC_SPEC_SELGROUP": "0"}, {"MPC_SPEC_PM2": "0"}])
def test_generation_variants_compile_with_nvrtc(mpcb, monkeypatch, flags):
    """The generation-time switches of the specialised kernel (regrouping queues, fused / adaptive / shared winner pass, deferral buffer,
    selector groups off, first-generation plane-major form)."""
    for k, v in flags.items():
        monkeypatch.setenv(k, v)
    for cfg in ("F4", "P6"):
        rc, nbytes, log = mpcb.jit_compile_check(cfg_path(cfg))
        assert rc == 0 and nbytes > 10000, log


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 3, 5])
def test_jit_kernels_with_regrouping_queues_match_oracle(mpcb, monkeypatch, seed):
    """MPC_SPEC_REGROUP=1 (off by default: measured slower, profiles/r02_regroup.txt): warps whose lanes picked different
    winners hand their blocks to per-module shared-memory queues; full batches of one module are taken by any warp, the
    last warps drain the partial ones.  Finely mixed data keeps the queues busy; ragged counts exercise the drain."""
    from tools.gen_dump import synth
    monkeypatch.setenv("MPC_SPEC_REGROUP", "1")
    monkeypatch.setenv("MPC_JIT_CACHE_DIR", "")
    cfg = eligible_config(300 + seed)
    while sum(1 for m in cfg["modules"].values() if m["name"] == "PredComp") < 2:  # queues need two modules to choose from
        seed += 17
        cfg = eligible_config(300 + seed)
    m = mpcb.Mpc(mpcb.load_config(text=json.dumps(cfg)))
    assert m.kernel_name() == "spec_thread:jit"
    rng = np.random.default_rng(seed)
    for blocks in (random_blocks(rng, 33), random_blocks(rng, 40001), synth("mixed_hashed", 7, 0, 300007, 300007)):
        sizes, sels, st = m.compress(blocks)
        r = OracleMPC(cfg).run(blocks)
        bad = np.nonzero((sizes != r.sizes) | (sels != r.sels))[0]
        assert bad.size == 0, (len(blocks), bad[:5], sizes[bad[:5]], r.sizes[bad[:5]])
        assert st.CompressedSize == r.CompressedSize and np.array_equal(st.count, r.count)
        assert np.array_equal(st.res_abs, r.res_abs) and np.array_equal(st.res_sq, r.res_sq)
