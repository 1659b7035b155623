"""PATTERN, the reference's data-pattern analysis (Pattern.cpp, Pattern.h, LRU.h): the oracle against the unmodified
reference on the CPU, the device block function on the host, the GPU path against the oracle through the C ABI, and the
CLI's CSV against the reference binary's."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, random_blocks
from oracle.bridge import PATTERN_WORDS, REF_BIN, RefCompressor, have_ref, oracle_pattern
from tools.gen_dump import KINDS, kat_blocks, synth

KAT = [204, 204, 316, 1028, 1028, 324, 260, 316, 204]  # SURVEY.md section 8c, "PATTERN ret"
BASE = [8, 8, 8, 4, 4, 2]


def dump_with_repeats(seed=5, n=6000):
    """mixed classes with whole-line repeats at several distances (temporal locality) and the known-answer blocks"""
    rng = np.random.default_rng(seed)
    d = np.concatenate([kat_blocks(), synth("mixed_hashed", seed, 0, n, n), kat_blocks()])
    idx = rng.integers(0, d.shape[0], d.shape[0] // 3)
    return np.concatenate([d, d[idx], random_blocks(rng, 500), d[idx[::-1]]])


GOLD = os.path.join(ROOT, "tests", "golden")


def golden_case():
    """the dump of tests/golden/make_golden_pattern.py and what the unmodified reference produced for it"""
    sys_path_golden = os.path.join(ROOT, "tests", "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_pattern", os.path.join(sys_path_golden, "make_golden_pattern.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(GOLD, "pattern.npz"))
    d = mod.pattern_dump()
    assert d.shape[0] == int(g["n"][0])
    return d, g["sizes"].astype(np.uint32), g["stats"]


def test_oracle_matches_committed_reference_fixture():
    d, sizes, stats = golden_case()
    got_sizes, got_stats = oracle_pattern(d)
    assert np.array_equal(got_sizes, sizes) and np.array_equal(got_stats, stats)


def test_oracle_known_answers():
    sizes, st = oracle_pattern(kat_blocks())
    assert sizes.tolist() == KAT
    assert st[4] == 9 * 128 and st[0] == 128 and st[1] == 3 * 128 and st[2] == 0


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("L", [32, 64, 128])
def test_oracle_matches_reference(L):
    rng = np.random.default_rng(L)
    d = random_blocks(rng, 1500, L)
    d = np.concatenate([d, d[rng.integers(0, 1500, 700)], np.zeros((3, L), np.uint8), np.full((2, L), 0xAB, np.uint8)])
    if L == 128:
        d = np.concatenate([d, dump_with_repeats()])
    sizes, st = oracle_pattern(d, L)
    ref = RefCompressor("PATTERN", None, L)
    rs, _ = ref.compress(d)
    assert np.array_equal(sizes, rs)
    assert np.array_equal(st, ref.pattern_stats())
    assert ref.totals()[2] == 0  # the reference never calls Update: the CLI prints "comp.ratio: 0"


def test_oracle_fifo_eviction_rule():
    """LRU.h:17-56 as Pattern.cpp:101-107 drives it: hits do not refresh, misses insert, the oldest insert is dropped."""
    def line(v):
        return np.full(128, v, np.uint8)
    seq = [1, 2, 3, 1, 4, 1, 2, 2, 5, 3]
    d = np.stack([line(v) for v in seq])
    # capacity 3: 1 2 3 miss | 1 hit | 4 miss (evicts 1) | 1 miss (evicts 2) | 2 miss (evicts 3) | 2 hit | 5 miss (evicts 4) | 3 miss
    _, st = oracle_pattern(d, 128, capacity=3)
    assert st[2] == 2 * 128
    _, st = oracle_pattern(d, 128)
    assert st[2] == 5 * 128


@pytest.fixture(scope="module")
def swar(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("swarp") / "libswar.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "cal_22-mpc_b200", "csrc"),
                    os.path.join(ROOT, "tests", "swar_host.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def test_block_function_matches_oracle(swar):
    d = np.ascontiguousarray(dump_with_repeats(11, 5000))
    n = d.shape[0]
    sizes, sels, imms = np.zeros(n, np.uint32), np.zeros(n, np.int32), np.zeros(n, np.uint32)
    hashes = np.zeros(n, np.uint64)
    swar.t_pattern_run.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong] + [ctypes.c_void_p] * 4
    swar.t_pattern_run(d.ctypes.data, n, sizes.ctypes.data, sels.ctypes.data, imms.ctypes.data, hashes.ctypes.data)
    want, st = oracle_pattern(d)
    assert np.array_equal(sizes, want)
    for p in range(6):
        m = sels == p
        assert int(imms[m].sum()) * BASE[p] == st[5 + p]
        assert (int(m.sum()) * (128 // BASE[p]) - int(imms[m].sum())) * BASE[p] == st[11 + p]
    assert int((sels == 9).sum()) * 128 == st[3]
    # equal lines hash equally, and the hash separates the distinct lines of this dump
    _, first = np.unique(d, axis=0, return_index=True)
    assert len(np.unique(hashes)) == len(first)


def check_against_oracle(mpcb, d, cache_blocks=0):
    sizes, st, _ = mpcb.pattern_run(d, cache_blocks=cache_blocks)
    want, words = oracle_pattern(d, 128, cache_blocks or (1 << 24) - 1) if d.shape[0] else (np.zeros(0, np.uint32), np.zeros(PATTERN_WORDS, np.uint64))
    assert np.array_equal(sizes.astype(np.uint32), want)
    assert np.array_equal(st.words(), words)
    assert st.blocks == d.shape[0]
    return st


@pytest.mark.gpu
@pytest.mark.parametrize("L", [32, 64])
def test_gpu_short_lines(mpcb, L):
    """32- / 64-byte lines: four / two lines per thread, duplicates confirmed on L bytes; ragged counts, a cache that evicts"""
    rng = np.random.default_rng(L)
    for d, cache in ((np.concatenate([dump_with_repeats().reshape(-1, L), random_blocks(rng, 1501, L)]), 0),
                     (synth("mixed_hashed", 5, 0, 900, 900).reshape(-1, L)[:-3], 0),
                     (np.concatenate([random_blocks(rng, 300, L)] * 3), 100)):
        sizes, st, _ = mpcb.pattern_run(d, cache_blocks=cache, line_size=L)
        want, words = oracle_pattern(d, L, cache or (1 << 24) - 1)
        assert np.array_equal(sizes.astype(np.uint32), want)
        assert np.array_equal(st.words(), words)
        assert st.blocks == d.shape[0]


@pytest.mark.gpu
def test_gpu_known_answers_and_classes(mpcb):
    sizes, st, _ = mpcb.pattern_run(kat_blocks())
    assert sizes.tolist() == KAT
    for kind in KINDS:
        check_against_oracle(mpcb, synth(kind, 31, 999, 3001, 1 << 20))
    st = check_against_oracle(mpcb, dump_with_repeats())
    assert st.temporal_path == 0 and st.temporal_bytes > 0
    assert st.distinct_blocks == len(np.unique(dump_with_repeats(), axis=0))


@pytest.mark.gpu
def test_gpu_matches_committed_reference_fixture(mpcb, tmp_path):
    """library and CLI against tests/golden (generated from the unmodified reference): per-line values, counters, CSV bytes"""
    d, sizes, stats = golden_case()
    got, st, _ = mpcb.pattern_run(d)
    assert np.array_equal(got.astype(np.uint32), sizes) and np.array_equal(st.words(), stats)
    ds = tmp_path / "ds"
    out = tmp_path / "out"
    ds.mkdir()
    out.mkdir()
    np.save(ds / "pattern_set.npy", np.concatenate([d, np.zeros((1, 128), np.uint8)]))
    for _ in range(2):  # appended twice, like the fixture: header + 2 rows
        r = subprocess.run([os.path.join(ROOT, "bin", "compressor"), "-a", "PATTERN", "-i", str(ds / "pattern_set.npy"), "-o", str(out)],
                           capture_output=True, text=True, check=True)
    assert r.stdout == open(os.path.join(GOLD, "cli_PATTERN_stdout.txt")).read()
    assert (out / "PATTERN_results.csv").read_bytes() == open(os.path.join(GOLD, "cli_PATTERN_results.csv"), "rb").read()
    assert not (out / "PATTERN_results_detail.csv").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 31, 33])
def test_gpu_ragged(mpcb, n):
    rng = np.random.default_rng(n)
    check_against_oracle(mpcb, random_blocks(rng, n) if n else np.zeros((0, 128), np.uint8))


@pytest.mark.gpu
def test_gpu_small_cache_takes_the_in_order_path(mpcb):
    """more distinct lines than the cache holds: evictions make the count order-dependent -> exact host pass"""
    d = dump_with_repeats(3, 3000)
    st = check_against_oracle(mpcb, d, cache_blocks=100)
    assert st.temporal_path == 1
    st = check_against_oracle(mpcb, d, cache_blocks=1 << 20)
    assert st.temporal_path == 0


@pytest.mark.gpu
def test_gpu_device_pointer_at_scale(mpcb):
    """256 MiB resident mixed dump: counters of two halves add up to the whole except temporal locality, which can only
    grow when the halves see each other; the whole equals the oracle on a window."""
    import torch
    n = (256 << 20) // 128
    m = mpcb.Mpc(__import__("helpers").cfg_path("P6"))
    d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    m.synth_device(d.data_ptr(), 0, n, n, "mixed_hashed", 77)
    _, whole, ms = mpcb.pattern_run(device_ptr=d.data_ptr(), n_blocks=n)
    _, a, _ = mpcb.pattern_run(device_ptr=d.data_ptr(), n_blocks=n // 2)
    _, b, _ = mpcb.pattern_run(device_ptr=d.data_ptr() + (n // 2) * 128, n_blocks=n - n // 2)
    wa, wb, ww = a.words(), b.words(), whole.words()
    keep = np.ones(PATTERN_WORDS, bool)
    keep[2] = False
    assert np.array_equal((wa + wb)[keep], ww[keep])
    assert ww[2] >= wa[2] + wb[2]
    win = d[: 20000 * 128].cpu().numpy().reshape(-1, 128)
    check_against_oracle(mpcb, win)
    m.close()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built")
def test_cli_csv_equals_reference(tmp_path):
    d = dump_with_repeats(9, 4000)
    ds = tmp_path / "bench"
    ds.mkdir()
    np.save(ds / "app.npy", np.concatenate([d, d[:1]]))  # the loader drops the last row
    outs = []
    for exe, name in ((REF_BIN, "ref"), (os.path.join(ROOT, "bin", "compressor"), "ours")):
        out = tmp_path / name
        out.mkdir()
        r = subprocess.run([exe, "-a", "PATTERN", "-i", str(ds / "app.npy"), "-o", str(out)], capture_output=True, text=True, check=True)
        outs.append((r.stdout, (out / "PATTERN_results.csv").read_bytes(), os.path.exists(out / "PATTERN_results_detail.csv")))
    assert outs[0] == outs[1]
    assert outs[0][0].strip() == "comp.ratio: 0"
