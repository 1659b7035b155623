"""Stateless secondary compressors (BASELINE.json config #5: BDI, FPC, BPC): oracle vs the unmodified reference on
the CPU, GPU kernels vs the oracle through the C ABI."""
import numpy as np
import pytest

from helpers import random_blocks
from oracle.bridge import RefCompressor, have_ref, oracle_variant
from tools.gen_dump import KINDS, kat_blocks, synth

ALGS = ["BDI", "FPC", "BPC"]
# SURVEY.md section 8c known answers (per-line return value for the nine known-answer blocks)
KAT = {"BDI": [12, 68, 316, 1028, 1028, 324, 260, 316, 68],
       "FPC": [6, 1120, 319, 1120, 1104, 848, 32, 149, 224],
       "BPC": [14, 14, 19, 1063, 93, 368, 98, 68, 14]}


@pytest.mark.parametrize("alg", ALGS)
def test_oracle_known_answers(alg):
    sizes, _ = oracle_variant(alg, kat_blocks())
    assert sizes.tolist() == KAT[alg]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("L", [32, 64, 128])
def test_oracle_matches_reference(alg, L):
    rng = np.random.default_rng(L)
    d = random_blocks(rng, 2500, L)
    if L == 128:
        d = np.concatenate([d, synth("mixed_hashed", 9, 0, 4000, 4000)])
    sizes, counts = oracle_variant(alg, d, L)
    ref = RefCompressor(alg, None, L)
    rs, _ = ref.compress(d)
    assert np.array_equal(sizes, rs)
    orig, comp, _ = ref.totals()
    assert comp == int(sizes.astype(np.uint64).sum())
    rc = ref.counts()
    if alg == "BDI":
        assert np.array_equal(rc, counts[:9]) and orig == d.shape[0] * 8 * L
    elif alg == "BPC":
        assert rc[0] == counts[7] and np.array_equal(rc[1:8], counts[:7])
    else:  # FPC: the reference's statistics can be inflated by its read past the line end (FPC.cpp:26)
        assert rc[0] >= counts[:8].sum() and np.all(rc[2:9] == counts[1:8]) and rc[1] >= counts[0]


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("L", [32, 64])
def test_gpu_short_lines(mpcb, alg, L):
    """32- / 64-byte lines (GPGPU-Sim sectors): four / two lines per thread; ragged counts so that the last 128-byte unit is partial"""
    rng = np.random.default_rng(L)
    for n in (1, 3, 5, 4099):
        d = random_blocks(rng, n, L)
        if n > 1000:
            d = np.concatenate([d, synth("mixed_hashed", 5, 0, 3000, 3000).reshape(-1, L)[:-1]])
        sizes, st, _ = mpcb.variant_run(alg, d, line_size=L)
        want, counts = oracle_variant(alg, d, L)
        assert np.array_equal(sizes.astype(np.uint32), want), n
        assert st.blocks == d.shape[0] and st.original_bits == d.shape[0] * 8 * L
        assert st.compressed_bits == int(want.astype(np.uint64).sum())
        assert np.array_equal(np.array(st.counts[:9], dtype=np.uint64), counts[:9]), n


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ALGS)
def test_gpu_known_answers_and_classes(mpcb, alg):
    sizes, st, _ = mpcb.variant_run(alg, kat_blocks())
    assert sizes.tolist() == KAT[alg]
    for kind in KINDS:
        d = synth(kind, 31, 999, 3001, 1 << 20)
        sizes, st, ms = mpcb.variant_run(alg, d)
        want, counts = oracle_variant(alg, d)
        assert np.array_equal(sizes.astype(np.uint32), want), kind
        assert st.blocks == 3001 and st.original_bits == 3001 * 1024
        assert st.compressed_bits == int(want.astype(np.uint64).sum())
        assert np.array_equal(np.array(st.counts[:9], dtype=np.uint64), counts[:9]), kind


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("n", [0, 1, 31, 33])
def test_gpu_ragged(mpcb, alg, n):
    rng = np.random.default_rng(n)
    d = random_blocks(rng, n) if n else np.zeros((0, 128), np.uint8)
    sizes, st, _ = mpcb.variant_run(alg, d)
    want, counts = oracle_variant(alg, d) if n else (np.zeros(0, np.uint32), np.zeros(16, np.uint64))
    assert np.array_equal(sizes.astype(np.uint32), want) and st.blocks == n
    assert np.array_equal(np.array(st.counts[:9], dtype=np.uint64), counts[:9])


@pytest.mark.gpu
def test_gpu_variants_at_scale_are_consistent(mpcb):
    """256 MiB resident dump: totals equal the sum of per-block sizes of a host run over a window, and the
    statistics of two disjoint halves add up to the whole (linearity)."""
    import torch
    n = (256 << 20) // 128
    m = mpcb.Mpc(__import__("helpers").cfg_path("P6"))
    d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    m.synth_device(d.data_ptr(), 0, n, n, "mixed_hashed", 5)
    m.sync()
    for alg in ALGS:
        _, whole, ms = mpcb.variant_run(alg, device_ptr=d.data_ptr(), n_blocks=n)
        _, a, _ = mpcb.variant_run(alg, device_ptr=d.data_ptr(), n_blocks=n // 2)
        _, b, _ = mpcb.variant_run(alg, device_ptr=d.data_ptr() + (n // 2) * 128, n_blocks=n - n // 2)
        assert whole.compressed_bits == a.compressed_bits + b.compressed_bits
        assert list(whole.counts) == [x + y for x, y in zip(a.counts, b.counts)]
        w0 = 100000
        want, _ = oracle_variant(alg, synth("mixed_hashed", 5, w0, 5000, n))
        sizes, _, _ = mpcb.variant_run(alg, d[w0 * 128:(w0 + 5000) * 128].cpu().numpy())
        assert np.array_equal(sizes.astype(np.uint32), want)
        print(alg, "GB/s", n * 128 / ms / 1e6, "ratio", whole.original_bits / whole.compressed_bits)


# ---- CPACK (host, sequential) and SC2 (GPU histogram + host tree + GPU lookup) -----------------------------------------
KAT_CPACK = [64, 220, 1056, 1088, 1088, 668, 128, 516, 192]  # SURVEY.md section 8c, this order in one process


def _mixed(n, seed=3):
    rng = np.random.default_rng(seed)
    return np.concatenate([kat_blocks(), synth("mixed_hashed", seed, 0, n, n), random_blocks(rng, n // 6)])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_cpack_sc2_oracles_match_reference():
    from oracle.bridge import oracle_cpack, oracle_sc2
    d = _mixed(30000)
    sizes, counts = oracle_cpack(d)
    assert sizes[:9].tolist() == KAT_CPACK
    rs, _ = RefCompressor("CPACK", None, 128).compress(d)
    assert np.array_equal(sizes, rs)
    for S in (10000, 20000, 300):
        ref = RefCompressor("SC2", None, 128, S)
        rs, _ = ref.compress(d)
        assert np.array_equal(oracle_sc2(d, S), rs), S
    few = synth("sparse_i32", 3, 0, 40000, 40000)  # fewer than 1024 distinct symbols: no trimming
    rs, _ = RefCompressor("SC2", None, 128, 10000).compress(few)
    assert np.array_equal(oracle_sc2(few, 10000), rs)
    for L in (32, 64):  # SC2 works word by word: the same dump bytes as shorter lines
        dl = d.reshape(-1, L)[:50001]
        rs, _ = RefCompressor("SC2", None, L, 12000).compress(dl)
        assert np.array_equal(oracle_sc2(dl, 12000, L), rs), L


def test_cpack_host_matches_oracle(mpcb):
    from oracle.bridge import oracle_cpack
    d = _mixed(20000)
    sizes, st = mpcb.cpack_run(d)
    want, counts = oracle_cpack(d)
    assert np.array_equal(sizes.astype(np.uint32), want)
    assert st.compressed_bits == int(want.astype(np.uint64).sum()) and st.blocks == d.shape[0]
    assert np.array_equal(np.array(st.counts[:6], dtype=np.uint64), counts[:6])


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,S", [("mixed_hashed", 60000, 10000), ("mixed_hashed", 60000, 25000), ("sparse_i32", 40000, 10000),
                                      ("random", 30000, 10000), ("zero", 20000, 10000), ("mixed_hashed", 5000, 10000)])
def test_sc2_gpu_matches_oracle(mpcb, kind, n, S):
    from oracle.bridge import oracle_sc2
    d = synth(kind, 17, 0, n, n)
    sizes, st, ms = mpcb.sc2_run(d, S)
    want = oracle_sc2(d, S)
    bad = np.nonzero(sizes.astype(np.uint32) != want)[0]
    assert bad.size == 0, (bad[:5], sizes[bad[:5]], want[bad[:5]])
    assert st.compressed_bits == int(want.astype(np.uint64).sum()) and st.original_bits == n * 1024
    assert mpcb.sc2_sampling_lines(n + 1) == max(10000, min((n + 1) // 100, 1000000))  # main.cpp:108-114


@pytest.mark.gpu
@pytest.mark.parametrize("L", [32, 64])
def test_sc2_gpu_short_lines(mpcb, L):
    from oracle.bridge import oracle_sc2
    d = synth("mixed_hashed", 23, 0, 20000, 20000).reshape(-1, L)[:-3]  # ragged: the last 128-byte unit is partial
    for S in (10000, 30000):
        sizes, st, _ = mpcb.sc2_run(d, S, line_size=L)
        want = oracle_sc2(d, S, L)
        bad = np.nonzero(sizes.astype(np.uint32) != want)[0]
        assert bad.size == 0, (L, S, bad[:5], sizes[bad[:5]], want[bad[:5]])
        assert st.compressed_bits == int(want.astype(np.uint64).sum()) and st.original_bits == d.shape[0] * 8 * L


@pytest.mark.gpu
def test_sc2_two_phase_api_shards_like_one_pass(mpcb):
    """SURVEY.md section 8e for SC2: the table comes from the shard that holds the sampling window, every shard applies it to
    its own lines with its global offset -- per-line sizes and totals equal the one-pass result and the CPU oracle."""
    import ctypes as C
    import torch
    from oracle.bridge import oracle_sc2
    from tools.gen_dump import synth
    n, S = 30000, 10000
    blocks = synth("mixed_hashed", 99, 0, n, n)
    want = oracle_sc2(blocks, S)
    d = torch.from_numpy(blocks).cuda()
    lib = mpcb.lib()
    table = mpcb.capi.Sc2Table()
    assert lib.mpc_sc2_build_table(0, d.data_ptr(), S, 128, C.byref(table)) == 0, lib.mpc_sc2_error()
    assert 0 < table.n <= 1024
    got = np.zeros(n, np.uint16)
    total = 0
    for lo, hi in ((0, 12345), (12345, n)):  # two "ranks"
        sizes = torch.zeros(hi - lo, dtype=torch.int16, device="cuda")
        st, ms = mpcb.VariantStats(), C.c_float()
        rc = lib.mpc_sc2_apply_device(0, d.data_ptr() + lo * 128, hi - lo, lo, S, 128, C.byref(table), sizes.data_ptr(), C.byref(st), C.byref(ms))
        assert rc == 0, lib.mpc_sc2_error()
        got[lo:hi] = sizes.cpu().numpy().view(np.uint16)
        total += st.compressed_bits
    assert np.array_equal(got.astype(np.uint32), want)
    assert total == int(want.astype(np.uint64).sum())
    one, st1, _ = mpcb.sc2_run(blocks, S)
    assert np.array_equal(one.astype(np.uint32), want) and st1.compressed_bits == total
