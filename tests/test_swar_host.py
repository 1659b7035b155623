"""Exhaustive CPU checks of the bit tricks the kernels are built from (csrc/mpc_device.cuh compiled with g++)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT


@pytest.fixture(scope="module")
def swar(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("swar") / "libswar.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "cal_22-mpc_b200", "csrc"),
                    os.path.join(ROOT, "tests", "swar_host.cpp"), "-o", so], check=True)
    l = ctypes.CDLL(so)
    l.t_zero_run_cost.argtypes = [ctypes.c_ulonglong]
    l.t_lzr.argtypes = [ctypes.c_ulonglong, ctypes.c_uint]
    for name in ("t_sub", "t_add", "t_xc", "t_xf", "t_zero_run_cost", "t_lzr", "t_row2_cost"):
        getattr(l, name).restype = ctypes.c_uint
    return l


def row_cost_reference(v):
    """FPCModule.cpp:47-66 on a 16-bit row whose bit 15 is scan position 0."""
    bits = [(v >> (15 - k)) & 1 for k in range(16)]
    ones = [k for k in range(16) if bits[k]]
    if not ones:
        return 0
    if len(ones) == 1:
        return 7
    if len(ones) == 2 and ones[1] - ones[0] == 1:
        return 8
    if not any(bits[:8]) or not any(bits[8:]):
        return 12
    return 17


def test_row_cost_exhaustive(swar):
    want = np.array([row_cost_reference(v) for v in range(65536)], dtype=np.uint32)
    lo = np.zeros(65536, np.uint32)
    hi = np.zeros(65536, np.uint32)
    for other in (0x0000, 0xFFFF, 0x0001, 0x8000, 0x8001, 0x0180, 0x5A5A):
        swar.t_row_costs(other, lo.ctypes.data, hi.ctypes.data)
        for arr in (lo, hi):
            assert np.array_equal(arr & 0x7FFFFFFF, want), hex(other)
            assert np.array_equal(arr >> 31, (np.arange(65536) != 0).astype(np.uint32)), hex(other)


def test_bytewise_add_sub_xor(swar):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2**32, 20000, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2**32, 20000, dtype=np.uint64).astype(np.uint32)
    a[:8] = [0, 0xFFFFFFFF, 0x80808080, 0x7F7F7F7F, 0x00FF00FF, 0x01010101, 0xFF00FF00, 0x80000000]
    b[:8] = [0xFFFFFFFF, 0, 0x01010101, 0x80808080, 0xFF00FF00, 0x02020202, 0x00FF00FF, 0x00000080]
    ab, bb = a.view(np.uint8).reshape(-1, 4), b.view(np.uint8).reshape(-1, 4)
    sub = (ab.astype(np.int32) - bb).astype(np.uint8).view(np.uint32).reshape(-1)
    add = (ab.astype(np.int32) + bb).astype(np.uint8).view(np.uint32).reshape(-1)
    xc = (ab ^ (ab >> 1)).view(np.uint32).reshape(-1)
    xf = (ab ^ np.where(ab & 0x80, 0x7F, 0).astype(np.uint8)).view(np.uint32).reshape(-1)
    for i in range(a.size):
        assert swar.t_sub(int(a[i]), int(b[i])) == int(sub[i])
        assert swar.t_add(int(a[i]), int(b[i])) == int(add[i])
        assert swar.t_xc(int(a[i]), 0) == int(xc[i]) and swar.t_xf(int(a[i]), 0) == int(xf[i])
        keep = swar.t_xc(int(a[i]), 0xFF)  # byte lane 0 untouched (column 0, XORModule.cpp:12)
        assert keep & 0xFF == int(a[i]) & 0xFF and keep >> 8 == int(xc[i]) >> 8
        keep = swar.t_xf(int(a[i]), 0xFF)
        assert keep & 0xFF == int(a[i]) & 0xFF and keep >> 8 == int(xf[i]) >> 8


def test_zero_runs_and_leading_zero_rows(swar):
    rng = np.random.default_rng(4)
    masks = [0, 2**64 - 1, 1, 2**63, 0b1011, 0x5555555555555555, 0xAAAAAAAAAAAAAAAA] + \
        [int(v) for v in rng.integers(0, 2**63, 3000)] + [int(v) & int(w) for v, w in zip(rng.integers(0, 2**63, 2000), rng.integers(0, 2**63, 2000))]
    for z in masks:
        cost, run = 0, 0
        for i in range(64):  # FPCModule.cpp:27-45, 69-79
            if (z >> i) & 1:
                run += 1
            elif run:
                cost += 7 if run > 1 else 4
                run = 0
        if run:
            cost += 7 if run > 1 else 4
        assert swar.t_zero_run_cost(z) == cost, hex(z)
        lz = 0
        while lz < 64 and (z >> lz) & 1:
            lz += 1
        assert swar.t_lzr(z, 64) == lz
        assert swar.t_lzr(z & 0xFFFF, 16) == min(lz, 16)


@pytest.mark.parametrize("alg", ["BDI", "FPC", "BPC"])
def test_variant_block_functions_match_oracle(swar, alg):
    """csrc/mpc_variants.cuh (the code the GPU runs per thread) compiled for the host vs the C oracle."""
    from helpers import random_blocks
    from oracle.bridge import VARIANT_ID, oracle_variant
    from tools.gen_dump import kat_blocks, synth
    rng = np.random.default_rng(8)
    d = np.concatenate([kat_blocks(), synth("mixed_hashed", 3, 0, 4000, 4000), random_blocks(rng, 3000)])
    want_sizes, want_counts = oracle_variant(alg, d)
    sizes = np.zeros(d.shape[0], np.uint32)
    counts = np.zeros(16, np.uint64)
    swar.t_variant_run.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
    swar.t_variant_run(VARIANT_ID[alg], d.ctypes.data, d.shape[0], sizes.ctypes.data, counts.ctypes.data)
    bad = np.nonzero(sizes != want_sizes)[0]
    assert bad.size == 0, (bad[:5], sizes[bad[:5]], want_sizes[bad[:5]])
    assert np.array_equal(counts, want_counts)


def test_bdi_range_tests_equal_the_reduce_sign_rule(swar):
    """bdi_fits64 / bdi_delta_fits32 (closed form) against reduceSign (BDI.cpp:203-218) on every boundary."""
    swar.t_bdi_fits.argtypes = swar.t_bdi_fits_rule.argtypes = [ctypes.c_ulonglong, ctypes.c_int]
    swar.t_bdi_delta32.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_int]
    M = (1 << 64) - 1
    rng = np.random.default_rng(11)
    for D in (1, 2, 4):
        half, full = 1 << (8 * D - 1), 1 << (8 * D)
        pts = {0, 1, half - 1, half, half + 1, full - 2, full - 1, full, full + 1, 1 << 62, (1 << 63) - 1, 1 << 63, (1 << 63) + 1}
        for k in (1, 2, 3, half - 1, half, half + 1, full - 1, full, full + 1, 1 << 40, 1 << 62):
            pts.add((-k) & M)
        for bits in range(1, 64):
            for dlt in (-1, 0, 1):
                pts.add(((1 << bits) + dlt) & M)
                pts.add((-(1 << bits) + dlt) & M)
        pts.update(int(v) for v in rng.integers(0, 1 << 63, 2000, dtype=np.uint64) * 2 + rng.integers(0, 2, 2000, dtype=np.uint64))
        for x in pts:
            assert swar.t_bdi_fits(x, D) == swar.t_bdi_fits_rule(x, D), (hex(x), D)
    for D in (1, 2):
        half, full = 1 << (8 * D - 1), 1 << (8 * D)
        bases = [0, 1, 5, half, full - 1, full, 0x7fffffff, 0x80000000, 0xffffffff - 3, 0xffffffff] + [int(v) for v in rng.integers(0, 1 << 32, 60)]
        for base in bases:
            for off in (0, 1, 2, 3, half - 1, half, half + 1, full - 1, full, full + 1, 0x7fffffff, 0x80000000, 0xfffffff0):
                for v in {(base + off) & 0xffffffff, (base - off) & 0xffffffff, off, 0xffffffff - off}:
                    assert swar.t_bdi_delta32(base, v, D) == swar.t_bdi_fits_rule((base - v) & M, D), (base, v, D)


@pytest.mark.parametrize("alg", ["BDI", "FPC", "BPC"])
@pytest.mark.parametrize("L", [32, 64, 128])
def test_variant_block_functions_for_every_line_size(swar, alg, L):
    """the W-word forms the GPU runs for 32- / 64- / 128-byte lines vs the C oracle (pinned to the reference at all three sizes)"""
    from helpers import random_blocks
    from oracle.bridge import VARIANT_ID, oracle_variant
    from tools.gen_dump import synth
    rng = np.random.default_rng(L + 1)
    d = np.concatenate([synth("mixed_hashed", 4, 0, 3000, 3000).reshape(-1, L), random_blocks(rng, 4000, L)])
    want_sizes, want_counts = oracle_variant(alg, d, L)
    sizes = np.zeros(d.shape[0], np.uint32)
    counts = np.zeros(16, np.uint64)
    swar.t_variant_run_l.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
    swar.t_variant_run_l(VARIANT_ID[alg], d.ctypes.data, d.shape[0], L, sizes.ctypes.data, counts.ctypes.data)
    bad = np.nonzero(sizes != want_sizes)[0]
    assert bad.size == 0, (bad[:5], sizes[bad[:5]], want_sizes[bad[:5]])
    assert np.array_equal(counts, want_counts)


def test_bdi_blocks_on_delta_boundaries_match_oracle(swar):
    """Blocks whose values sit on the immediate / delta limits of every (base size, delta size) pair."""
    from oracle.bridge import oracle_variant
    import random
    rnd = random.Random(12)
    blocks = []
    for B, dt in ((8, np.uint64), (4, np.uint32), (2, np.uint16)):
        n = 128 // B
        top = (1 << (8 * B)) - 1
        for D in (1, 2, 4):
            if D >= B:
                continue
            half, full = 1 << (8 * D - 1), 1 << (8 * D)
            for base in (full, full + 7, top - full, top, top - 1, (top >> 1) + 1, rnd.randrange(full, top)):
                for offs in ([0, -1, -2], [1, 2, half - 1, half, half + 1], [-half + 1, -half, -half - 1], [full - 1, full, -full],
                             [0] * 3, [-1] * 3, [half] * 3):
                    vals = np.full(n, base, dtype=object)
                    for j, o in enumerate(offs):
                        vals[1 + 3 * j % (n - 1)] = (base - o) & top
                    vals[n - 1] = rnd.randrange(0, full)  # an immediate
                    blocks.append(np.array([int(v) for v in vals], dtype=dt).view(np.uint8))
                    vals[0] = rnd.randrange(0, full)  # immediate first: the base is the second value
                    blocks.append(np.array([int(v) for v in vals], dtype=dt).view(np.uint8))
    d = np.stack(blocks)
    want_sizes, want_counts = oracle_variant("BDI", d)
    sizes = np.zeros(d.shape[0], np.uint32)
    counts = np.zeros(16, np.uint64)
    swar.t_variant_run.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
    swar.t_variant_run(1, d.ctypes.data, d.shape[0], sizes.ctypes.data, counts.ctypes.data)
    bad = np.nonzero(sizes != want_sizes)[0]
    assert bad.size == 0, (bad[:5], sizes[bad[:5]], want_sizes[bad[:5]])
    assert np.array_equal(counts, want_counts)
    assert len(set(sizes.tolist())) > 8  # the set really exercises several encodings


def test_fpc_words_on_pattern_boundaries_match_oracle(swar):
    """fpc_block selects the pattern with range tests ((v + 2^(k-1)) < 2^k, a packed halfword add) instead of the reference's
    mask comparisons (FPC.cpp:16-83): every word on either side of every pattern boundary, alone in a line of zeros, after a
    zero and after a non-zero word, and mixed into random lines."""
    from oracle.bridge import VARIANT_ID, oracle_variant
    M = 0xFFFFFFFF
    words = {0, 1, 0x01010101, 0x7F7F7F7F, 0x80808080, 0xFFFFFFFF, 0x00FF00FF, 0xFF00FF00, 0x00010000, 0xFFFF0000, 0x0000FFFF}
    for k in (3, 7, 15, 16, 23, 31):  # sign-extension limits of 4 / 8 / 16 bits and the halfword / byte seams
        for d in (-2, -1, 0, 1, 2):
            words.add(((1 << k) + d) & M)
            words.add((-(1 << k) + d) & M)
    for hi in (0x0000, 0x007F, 0x0080, 0x00FF, 0xFF7F, 0xFF80, 0xFFFF, 0x8000, 0x7FFF, 0x0100):  # two sign-extended bytes or not
        for lo in (0x0000, 0x007F, 0x0080, 0x00FF, 0xFF7F, 0xFF80, 0xFFFF, 0x8000, 0x7FFF, 0x0100):
            words.add((hi << 16) | lo)
    for b in (0x00, 0x01, 0x7F, 0x80, 0xFE, 0xFF):  # one byte four times, and one byte off
        words.add(b * 0x01010101)
        words.add((b * 0x01010101) ^ 0x00010000)
    words = sorted(words)
    rng = np.random.default_rng(5)
    lines = []
    for w in words:
        a = np.zeros(32, np.uint32); a[5] = w; lines.append(a)                      # alone: zero runs on both sides
        a = np.zeros(32, np.uint32); a[0] = w; a[1] = w; a[31] = w; lines.append(a)  # first / repeated / last word
        a = rng.integers(0, 1 << 32, 32, dtype=np.uint64).astype(np.uint32); a[rng.integers(0, 32)] = w; lines.append(a)
    d = np.ascontiguousarray(np.stack(lines)).view(np.uint8).reshape(-1, 128)
    want_sizes, want_counts = oracle_variant("FPC", d)
    sizes = np.zeros(d.shape[0], np.uint32)
    counts = np.zeros(16, np.uint64)
    swar.t_variant_run.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
    swar.t_variant_run(VARIANT_ID["FPC"], d.ctypes.data, d.shape[0], sizes.ctypes.data, counts.ctypes.data)
    bad = np.nonzero(sizes != want_sizes)[0]
    assert bad.size == 0, (bad[:5], sizes[bad[:5]], want_sizes[bad[:5]])
    assert np.array_equal(counts, want_counts)
    assert all(want_counts[p] > 0 for p in range(8))  # every pattern occurs
    from oracle.bridge import RefCompressor, have_ref
    if have_ref():  # and the unmodified reference says the same on these lines
        ref_sizes, _ = RefCompressor("FPC").compress(d, want_sels=False)
        assert np.array_equal(ref_sizes, want_sizes)
