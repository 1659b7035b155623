"""GPU parity proper: the CUDA path, called through the C ABI (include/mpc_capi.h), against the CPU oracle and
the committed reference goldens.  Bit-exact: same selected cluster and same compressed bit count per block,
same totals, histograms and MAE/MSE numerators."""
import json

import numpy as np
import pytest

from helpers import SHIPPED, cfg_path, random_blocks, random_config
from oracle.bridge import OracleMPC
from tools.gen_dump import KINDS, kat_blocks, synth

pytestmark = pytest.mark.gpu

KERNELS = [1, 0]  # 1 = generic warp-per-block; 0 = auto (specialised when built in)


def check_against_oracle(mpcb, m, oracle, blocks):
    sizes, sels, st = m.compress(blocks)
    r = oracle.run(blocks)
    bad = np.nonzero((sizes != r.sizes) | (sels != r.sels))[0]
    assert bad.size == 0, f"{bad.size} blocks differ, first {bad[:5]}: gpu {sizes[bad[:5]]}/{sels[bad[:5]]} " \
                          f"oracle {r.sizes[bad[:5]]}/{r.sels[bad[:5]]} kernel {m.kernel_name()}"
    assert st.blocks == r.blocks and st.OriginalSize == r.OriginalSize and st.CompressedSize == r.CompressedSize
    assert np.array_equal(st.count, r.count) and np.array_equal(st.comp_bits, r.comp_bits)
    assert np.array_equal(st.res_lines, r.res_lines)
    assert np.array_equal(st.res_abs, r.res_abs) and np.array_equal(st.res_sq, r.res_sq)
    hb = r.hist.shape[1]
    assert np.array_equal(st.hist[:, :hb], r.hist) and not st.hist[:, hb:].any()
    return st


@pytest.mark.parametrize("kernel", KERNELS)
def test_known_answers(mpcb, kernel):
    m = mpcb.Mpc(cfg_path("P6"))
    m.set_kernel(kernel)
    sizes, sels, st = m.compress(kat_blocks())
    assert sizes.tolist() == [3, 35, 146, 1027, 1027, 1027, 201, 263, 35]
    assert sels.tolist() == [0, 1, 4, -1, -1, -1, 5, 3, 1]
    assert st.OriginalSize == 9 * 1024


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("cfg", SHIPPED)
def test_reference_golden(mpcb, golden, cfg, kernel):
    m = mpcb.Mpc(cfg_path(cfg))
    m.set_kernel(kernel)
    sizes, sels, st = m.compress(golden["blocks"])
    assert np.array_equal(sizes, golden[f"{cfg}_sizes"].astype(np.uint32))
    assert np.array_equal(sels, golden[f"{cfg}_sels"].astype(np.int32))
    orig, comp = (int(v) for v in golden[f"{cfg}_totals"])
    assert (st.OriginalSize, st.CompressedSize) == (orig, comp)
    assert st.CompRatio == float(golden[f"{cfg}_ratio"][0])
    stat, fl = golden[f"{cfg}_stat"], golden[f"{cfg}_fl"]
    assert np.array_equal(st.count, stat[:, 0]) and np.array_equal(st.comp_bits, stat[:, 2])
    for k in range(st.count.size):
        assert st.mae(k) == fl[k, 1] and st.mse(k) == fl[k, 2]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("cfg", SHIPPED)
@pytest.mark.parametrize("kind", KINDS)
def test_synthetic_classes(mpcb, cfg, kind, kernel):
    m = mpcb.Mpc(cfg_path(cfg))
    m.set_kernel(kernel)
    blocks = synth(kind, 4242, 12345, 3001, 1 << 20)  # ragged count on purpose
    check_against_oracle(mpcb, m, OracleMPC(cfg_path(cfg)), blocks)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("cfg,L", [("S32", 32), ("S64", 64)])
@pytest.mark.parametrize("kind", ["smooth_f32", "sparse_i32", "mixed_hashed", "wordsame"])
def test_short_line_shipped_configs(mpcb, cfg, L, kind, kernel):
    """32- / 64-byte lines (GPGPU-Sim sectors; the reference takes any lineSize, VPC.cpp:99-101): the same dump bytes cut into
    shorter lines, on the specialised kernel (four / two lines per thread) and on the generic one; ragged line count"""
    m = mpcb.Mpc(cfg_path(cfg))
    m.set_kernel(kernel)
    if kernel == 0:
        assert m.kernel_name() == "spec_thread:" + cfg
    blocks = synth(kind, 77, 999, 1501, 1 << 20).reshape(-1, L)[: 1501 * (128 // L) - 3]
    check_against_oracle(mpcb, m, OracleMPC(cfg_path(cfg)), blocks)


@pytest.mark.parametrize("seed", range(16))
def test_random_configs(mpcb, seed):
    rng = np.random.default_rng(7000 + seed)
    L = int(rng.choice([32, 64, 128]))
    cfg = random_config(rng, L=L)
    m = mpcb.Mpc(mpcb.load_config(text=json.dumps(cfg)))
    check_against_oracle(mpcb, m, OracleMPC(cfg), random_blocks(rng, 777, L))


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 255, 257])
def test_ragged_counts(mpcb, n):
    m = mpcb.Mpc(cfg_path("P6"))
    blocks = synth("mixed_hashed", 1, 0, n, max(n, 1))
    sizes, sels, st = m.compress(blocks)
    r = OracleMPC(cfg_path("P6")).run(blocks) if n else None
    assert st.blocks == n
    if n:
        assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels)
    else:
        assert st.CompressedSize == 0 and st.OriginalSize == 0


def test_statistics_accumulate_and_reset(mpcb):
    m = mpcb.Mpc(cfg_path("F4"))
    a = synth("mixed_hashed", 3, 0, 1000, 2000)
    b = synth("mixed_hashed", 3, 1000, 1000, 2000)
    m.reset()
    m.submit_host(a)
    m.submit_host(b)
    st = m.finish()
    r = OracleMPC(cfg_path("F4")).run(np.concatenate([a, b]))
    assert st.blocks == 2000 and st.CompressedSize == r.CompressedSize and np.array_equal(st.count, r.count)
    m.reset()
    assert m.finish().blocks == 0


@pytest.mark.parametrize("cfg", ["F4", "P6"])
def test_back_to_back_launches_of_any_size_share_one_tile_counter(mpcb, cfg):
    """The specialised kernel hands out its last tiles through an atomic counter that the last CTA of a launch puts back to
    zero: launches of very different sizes, back to back on one stream without a reset in between, must each see every
    block exactly once (per-block results) and add up in the statistics."""
    import torch
    m = mpcb.Mpc(cfg_path(cfg))
    sizes_n = [1, 5000, 33, 148 * 20 * 32 * 9 + 17, 7, 300000, 64, 148 * 20 * 32 + 1]
    total = sum(sizes_n)
    blocks = synth("mixed_hashed", 21, 0, total, total)
    d = torch.from_numpy(blocks).cuda()
    packed = torch.zeros(total, dtype=torch.int16, device="cuda")
    m.reset()
    off = 0
    for rep in range(2):  # second round: the same launches again, the counter has been through every size already
        off = 0
        for n in sizes_n:
            m.submit_device(d.data_ptr() + off * 128, n, packed.data_ptr() + off * 2)
            off += n
    st = m.finish()
    r = OracleMPC(cfg_path(cfg)).run(blocks)
    got_sizes, got_sels = mpcb.unpack(packed.cpu().numpy().view(np.uint16))
    assert np.array_equal(got_sizes, r.sizes) and np.array_equal(got_sels, r.sels)
    assert st.blocks == 2 * total and st.CompressedSize == 2 * r.CompressedSize and np.array_equal(st.count, 2 * r.count)


def test_device_submit_and_synth_match_numpy(mpcb):
    import torch
    m = mpcb.Mpc(cfg_path("P6"))
    n, total, first = 5000, 1 << 22, 777777
    for kind in KINDS:
        d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
        m.synth_device(d.data_ptr(), first, n, total, kind, 99)
        m.sync()
        assert np.array_equal(d.cpu().numpy().reshape(n, 128), synth(kind, 99, first, n, total)), kind
    blocks = synth("mixed_hashed", 99, first, n, total)
    d = torch.from_numpy(blocks).cuda()
    packed = torch.zeros(n, dtype=torch.int16, device="cuda")
    m.reset()
    m.submit_device(d.data_ptr(), n, packed.data_ptr())
    st = m.finish()
    sizes, sels = mpcb.unpack(packed.cpu().numpy().view(np.uint16))
    r = OracleMPC(cfg_path("P6")).run(blocks)
    assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels) and st.CompressedSize == r.CompressedSize
    ms, launches = m.last_timing()
    assert launches == 1 and ms > 0


def test_chunked_host_path_multiple_chunks(mpcb):
    # more chunks (32 MiB each) than staging slots, so that every slot is reused (LoaderNPY replacement path)
    m = mpcb.Mpc(cfg_path("F4"))
    n = (160 << 20) // 128
    blocks = synth("mixed_regions", 5, 0, n, n)
    sizes, sels, st = m.compress(blocks)
    r = OracleMPC(cfg_path("F4")).run(blocks)
    assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels)
    assert st.CompressedSize == r.CompressedSize and np.array_equal(st.hist[:, :r.hist.shape[1]], r.hist)


@pytest.mark.parametrize("cfg,kind", [("F4", "smooth_f32"), ("P6", "mixed_hashed")])
def test_full_size_properties(mpcb, cfg, kind):
    """BASELINE sizes (1 GiB): size-independent properties instead of a CPU run of the whole dump --
    (1) per-block results of a seeded window equal the oracle, (2) totals equal the sum of the per-block
    results, (3) two half submissions accumulate to the same statistics as one (linearity)."""
    import torch
    m = mpcb.Mpc(cfg_path(cfg))
    n = (1 << 30) // 128
    d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    packed = torch.zeros(n, dtype=torch.int16, device="cuda")
    m.synth_device(d.data_ptr(), 0, n, n, kind, 2024)
    m.reset()
    m.submit_device(d.data_ptr(), n, packed.data_ptr())
    st = m.finish()
    sizes, sels = mpcb.unpack(packed.cpu().numpy().view(np.uint16))
    assert st.blocks == n and st.OriginalSize == n * 1024
    assert st.CompressedSize == int(sizes.astype(np.uint64).sum())
    assert np.array_equal(st.count, np.bincount(sels + 1, minlength=st.count.size).astype(np.uint64))
    w0, wn = 3_000_000, 20000
    r = OracleMPC(cfg_path(cfg)).run(synth(kind, 2024, w0, wn, n))
    assert np.array_equal(sizes[w0:w0 + wn], r.sizes) and np.array_equal(sels[w0:w0 + wn], r.sels)
    m.reset()
    half = n // 2
    m.submit_device(d.data_ptr(), half, None)
    m.submit_device(d.data_ptr() + half * 128, n - half, None)
    st2 = m.finish()
    assert st2.CompressedSize == st.CompressedSize and np.array_equal(st2.hist, st.hist)
    assert np.array_equal(st2.res_abs, st.res_abs) and np.array_equal(st2.res_sq, st.res_sq)


@pytest.mark.parametrize("cfg", SHIPPED)
def test_shipped_configs_use_the_specialised_kernel(mpcb, cfg):
    m = mpcb.Mpc(cfg_path(cfg))
    assert m.kernel_name() == "spec_thread:" + cfg
    m.set_kernel(1)
    assert m.kernel_name() == "generic_warp"
    m.set_kernel(2)
    assert m.kernel_name() == "spec_thread:" + cfg


def test_ineligible_config_falls_back_to_generic_and_says_so(mpcb, monkeypatch):
    rng = np.random.default_rng(5)
    monkeypatch.setenv("MPC_SPEC_BITGATHER", "0")  # arbitrary scan tables left to the generic kernel
    m = mpcb.Mpc(mpcb.load_config(text=json.dumps(random_config(rng, L=64, n_pred=2, table="perm"))))
    assert m.kernel_name() == "generic_warp"
    with pytest.raises(mpcb.MpcError) as e:
        m.set_kernel(2)
    assert "not eligible" in str(e.value)


def test_mixed_dump_fires_every_branch(mpcb):
    """BASELINE config #3 in miniature: on the hashed mixed dump every cluster is selected (zero, word-same, each
    predictor, raw fallback), zero rows appear both isolated and in runs, and every row pattern of the common
    encoder occurs -- checked on the oracle's view of the same blocks, then the GPU must agree bit for bit."""
    n = 200000
    blocks = synth("mixed_hashed", 77, 0, n, n)
    for cfg in ("P6", "E5"):
        m = mpcb.Mpc(cfg_path(cfg))
        sizes, sels, st = m.compress(blocks)
        r = OracleMPC(cfg_path(cfg)).run(blocks)
        assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels)
        if cfg == "P6":
            assert np.all(st.count > 0), (cfg, st.count)             # every cluster incl. -1 (raw) is hit
        assert st.hist[0, 8 * 128 + int(m.cfg.enc_bits[0])] > 0       # raw fallback size
        comp = sizes[sels >= m.cfg.first_predcomp]
        assert comp.min() < 100 and comp.max() > 900                  # from a handful of rows to nearly raw


def test_four_gib_mixed_properties(mpcb):
    """BASELINE config #3 size (4 GiB, 33 554 432 blocks): totals = sum of per-block results, window = oracle."""
    import torch
    cfg = "F4"
    m = mpcb.Mpc(cfg_path(cfg))
    n = (4 << 30) // 128
    d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    packed = torch.zeros(n, dtype=torch.int16, device="cuda")
    m.synth_device(d.data_ptr(), 0, n, n, "mixed_hashed", 31337)
    m.reset()
    m.submit_device(d.data_ptr(), n, packed.data_ptr())
    st = m.finish()
    p = packed.view(torch.int32)  # two packed results per word: sum sizes on the device
    lo = (p & 0x7FF).to(torch.int64).sum().item() + ((p >> 16) & 0x7FF).to(torch.int64).sum().item()
    assert st.blocks == n and st.CompressedSize == lo
    w0, wn = 20_000_000, 30000
    sizes, sels = mpcb.unpack(packed[w0:w0 + wn].cpu().numpy().view(np.uint16))
    r = OracleMPC(cfg_path(cfg)).run(synth("mixed_hashed", 31337, w0, wn, n))
    assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels)


def _full_dump_parity(mpcb, cfgs, kind, gib, seed, piece_mib=256):
    """Every block of a BASELINE-size dump against the threaded CPU oracle: the dump is generated on the device, the
    packed per-block results and the data come back in pieces, the oracle runs on each piece; totals, per-cluster counts,
    histograms and MAE/MSE numerators of the whole dump must equal the oracle's sums."""
    import torch
    n = (gib << 30) // 128
    d = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    packed = torch.zeros(n, dtype=torch.int16, device="cuda")
    ms = {c: mpcb.Mpc(cfg_path(c)) for c in cfgs}
    next(iter(ms.values())).synth_device(d.data_ptr(), 0, n, n, kind, seed)
    next(iter(ms.values())).sync()
    piece = (piece_mib << 20) // 128
    host_pieces = None
    for c, m in ms.items():
        m.reset()
        m.submit_device(d.data_ptr(), n, packed.data_ptr())
        st = m.finish()
        got = packed.cpu().numpy().view(np.uint16)
        sizes, sels = mpcb.unpack(got)
        orc = OracleMPC(cfg_path(c))
        tot = None
        mismatches = 0
        for lo in range(0, n, piece):
            hi = min(n, lo + piece)
            blocks = d[lo * 128:hi * 128].cpu().numpy().reshape(-1, 128)
            r = orc.run(blocks)
            mismatches += int(np.count_nonzero((sizes[lo:hi] != r.sizes) | (sels[lo:hi] != r.sels)))
            part = [r.blocks, r.OriginalSize, r.CompressedSize, r.count, r.comp_bits, r.res_lines, r.res_abs, r.res_sq, r.hist]
            tot = part if tot is None else [a + b for a, b in zip(tot, part)]
        assert mismatches == 0, f"{c}/{kind}: {mismatches} of {n} blocks differ from the oracle"
        assert (st.blocks, st.OriginalSize, st.CompressedSize) == (tot[0], tot[1], tot[2])
        assert np.array_equal(st.count, tot[3]) and np.array_equal(st.comp_bits, tot[4]) and np.array_equal(st.res_lines, tot[5])
        assert np.array_equal(st.res_abs, tot[6]) and np.array_equal(st.res_sq, tot[7])
        hb = tot[8].shape[1]
        assert np.array_equal(st.hist[:, :hb], tot[8]) and not st.hist[:, hb:].any()
    del host_pieces


def test_full_1gib_smooth_every_block_equals_oracle(mpcb):
    """BASELINE configs[1] (the headline workload), all 8 388 608 blocks."""
    _full_dump_parity(mpcb, ["F4"], "smooth_f32", 1, 2024)


def test_full_4gib_mixed_every_block_equals_oracle(mpcb):
    """BASELINE configs[2] (4 GiB, every branch), all 33 554 432 blocks, column-major (F4) and plane-major (P6) kernels."""
    _full_dump_parity(mpcb, ["F4", "P6"], "mixed_hashed", 4, 31337)


def test_finish_allreduce_on_one_gpu(mpcb):
    """The multi-GPU finish on a one-GPU job: without a communicator the all-reduce is the copy, and with a one-rank NCCL
    communicator (mpc_comm_unique_id + mpc_comm_init_rank: NCCL resolved with dlopen, every return code checked) the
    ncclAllReduce of the statistics vector must leave it unchanged; local statistics keep accumulating afterwards."""
    blocks = synth("mixed_hashed", 11, 0, 50000, 50000)
    r = OracleMPC(cfg_path("F4")).run(blocks)
    m = mpcb.Mpc(cfg_path("F4"))
    m.reset()
    m.submit_host(blocks)
    st = m.finish_allreduce()
    assert st.CompressedSize == r.CompressedSize and np.array_equal(st.count, r.count)
    m.comm_init_rank(mpcb.Mpc.comm_unique_id(), 1, 0)
    st = m.finish_allreduce()
    assert st.CompressedSize == r.CompressedSize and np.array_equal(st.count, r.count) and np.array_equal(st.res_sq, r.res_sq)
    m.submit_host(blocks)
    st = m.finish_allreduce()
    assert st.CompressedSize == 2 * r.CompressedSize and np.array_equal(st.hist[:, :r.hist.shape[1]], 2 * r.hist)
    with pytest.raises(mpcb.MpcError):
        m.comm_init_rank(mpcb.Mpc.comm_unique_id(), 1, 0)  # a context has one communicator


def test_two_gpu_shards_allreduce_equals_oracle(mpcb):
    """SURVEY.md section 8e on hardware: two contexts on two GPUs, contiguous shards, mpc_comm_init_all +
    mpc_finish_allreduce -- the reduced statistics equal the oracle's over the whole dump, per-block results per shard."""
    import ctypes as C
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 400001
    blocks = synth("mixed_hashed", 12, 0, n, n)
    r = OracleMPC(cfg_path("P6")).run(blocks)
    ms = [mpcb.Mpc(cfg_path("P6"), device=g) for g in range(2)]
    arr = (C.c_void_p * 2)(*[m.h for m in ms])
    assert mpcb.lib().mpc_comm_init_all(arr, 2) == 0, mpcb.lib().mpc_global_error()
    per = (n + 1) // 2
    packed = [np.zeros(per, np.uint16), np.zeros(n - per, np.uint16)]
    ms[0].submit_host(blocks[:per], packed[0])
    ms[1].submit_host(blocks[per:], packed[1])
    pod = mpcb.capi.StatsPod()
    assert mpcb.lib().mpc_finish_allreduce(arr, 2, C.byref(pod)) == 0, mpcb.lib().mpc_last_error(ms[0].h)
    st = mpcb.capi.Stats(pod, ms[0].cfg.num_modules, 128)
    sizes, sels = mpcb.unpack(np.concatenate(packed))
    assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels)
    assert st.blocks == n and st.CompressedSize == r.CompressedSize and np.array_equal(st.count, r.count)
    assert np.array_equal(st.res_abs, r.res_abs) and np.array_equal(st.hist[:, :r.hist.shape[1]], r.hist)


@pytest.mark.parametrize("direct", [False, True])
def test_submit_file_reads_the_dump_itself(mpcb, tmp_path, direct):
    """mpc_submit_file: pread straight into the pinned staging ring (more chunks than staging slots, an
    unaligned data offset like an .npy header's, O_DIRECT on and off) -- per-block results and statistics equal the oracle's."""
    import os
    n = (200 << 20) // 128 + 777
    blocks = synth("mixed_regions", 9, 0, n, n)
    path = str(tmp_path / "dump.bin")
    header = b"\x93NUMPY-like header of an odd length....."  # 40 bytes: the data starts unaligned
    with open(path, "wb") as f:
        f.write(header)
        f.write(blocks.tobytes())
    flags = os.O_RDONLY | (os.O_DIRECT if direct else 0)
    try:
        fd = os.open(path, flags)
    except OSError:
        pytest.skip("O_DIRECT not supported by the file system under tmp_path")
    try:
        m = mpcb.Mpc(cfg_path("F4"))
        packed = np.zeros(n, np.uint16)
        m.reset()
        try:
            m.submit_file(fd, len(header), n, packed, direct_io=direct)
        except mpcb.MpcError as e:
            if direct and "short read" in str(e):
                pytest.skip("O_DIRECT reads refused by the file system under tmp_path")
            raise
        st = m.finish()
    finally:
        os.close(fd)
    r = OracleMPC(cfg_path("F4")).run(blocks)
    sizes, sels = mpcb.unpack(packed)
    assert np.array_equal(sizes, r.sizes) and np.array_equal(sels, r.sels)
    assert st.blocks == n and st.CompressedSize == r.CompressedSize and np.array_equal(st.res_sq, r.res_sq)
