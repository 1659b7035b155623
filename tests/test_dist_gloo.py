"""N > 1 host logic on CPU: two gloo ranks shard a dump, each builds the statistics vector of its shard (here
from the CPU oracle, standing in for a GPU), the vectors are all-reduced and expanded by the library's host
function -- the result must equal the single-process statistics of the whole dump."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, cfg_path


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def stats_words_from_oracle(mpcb, r):
    K, HB = mpcb.capi.MAX_MODULES + 1, mpcb.capi.HIST_BINS
    w = np.zeros(mpcb.capi.STATS_WORDS, dtype=np.int64)
    k = r.count.size
    w[0:k] = r.res_abs.astype(np.int64)
    w[K:K + k] = r.res_sq.astype(np.int64)
    hist = np.zeros((K, HB), dtype=np.int64)
    hist[:k, :r.hist.shape[1]] = r.hist.astype(np.int64)
    w[2 * K:] = hist.reshape(-1)
    return w


def _worker(rank, world, port, cfg, n, out_dir):
    import importlib
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpcb = importlib.import_module("cal_22-mpc_b200")
    from oracle.bridge import OracleMPC
    from tools.gen_dump import synth
    lo, hi = mpcb.shard_range(n, rank, world)
    blocks = synth("mixed_hashed", 77, lo, hi - lo, n)
    r = OracleMPC(cfg_path(cfg)).run(blocks, threads=2)
    t = torch.from_numpy(stats_words_from_oracle(mpcb, r))
    mpcb.allreduce_stats(t)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5001, 4096])
def test_two_rank_sharding_and_allreduce(mpcb, tmp_path, n):
    cfg = "P6"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, cfg, n, str(tmp_path)), nprocs=2, join=True)
    reduced = np.load(tmp_path / "reduced.npy").view(np.uint64)
    from oracle.bridge import OracleMPC
    from tools.gen_dump import synth
    pod = mpcb.load_config(path=cfg_path(cfg))
    import ctypes
    out = mpcb.StatsPod()
    assert mpcb.lib().mpc_stats_expand(ctypes.byref(pod), reduced.ctypes.data, reduced.size, ctypes.byref(out)) == 0
    st = mpcb.Stats(out, pod.num_modules, pod.line_size)
    whole = OracleMPC(cfg_path(cfg)).run(synth("mixed_hashed", 77, 0, n, n))
    assert st.blocks == n and st.CompressedSize == whole.CompressedSize and st.OriginalSize == whole.OriginalSize
    assert np.array_equal(st.count, whole.count) and np.array_equal(st.comp_bits, whole.comp_bits)
    assert np.array_equal(st.res_abs, whole.res_abs) and np.array_equal(st.res_sq, whole.res_sq)
    assert np.array_equal(st.hist[:, :whole.hist.shape[1]], whole.hist)


def test_shard_ranges_cover_everything(mpcb):
    for n in (0, 1, 7, 8, 9, 1000, 1 << 22):
        for world in (1, 2, 3, 4, 8):
            ranges = [mpcb.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b and c <= d
