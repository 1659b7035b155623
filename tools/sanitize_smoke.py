#!/usr/bin/env python3
"""Every kernel of the library once on small inputs (a few thousand blocks, ragged tails) -- the program run under
`compute-sanitizer --tool memcheck` on a B200; results are checked against the CPU oracle so that the run is also a
parity smoke.  usage: sanitize_smoke.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mpcb = importlib.import_module("cal_22-mpc_b200")
    from oracle.bridge import OracleMPC, oracle_pattern, oracle_sc2, oracle_variant
    from tools.gen_dump import kat_blocks, synth
    blocks = np.concatenate([kat_blocks(), synth("mixed_hashed", 3, 0, 2500, 2500), synth("smooth_f32", 3, 0, 1000, 1000)])[:3333]
    for cfg in ("P6", "F4", "E5", "Z1"):
        path = os.path.join(ROOT, "configs", cfg + ".json")
        want = OracleMPC(path).run(blocks)
        m = mpcb.Mpc(path, device=0)
        for kernel in (0, 1):
            m.set_kernel(kernel)
            sizes, sels, st = m.compress(blocks)
            assert np.array_equal(sizes, want.sizes) and np.array_equal(sels, want.sels), (cfg, m.kernel_name())
            print("ok", cfg, m.kernel_name(), st.CompRatio)
        m.close()
    for alg in ("BDI", "FPC", "BPC"):
        sizes, st, _ = mpcb.variant_run(alg, blocks)
        want, _ = oracle_variant(alg, blocks)
        assert np.array_equal(sizes.astype(np.uint32), want), alg
        print("ok", alg)
    big = np.concatenate([blocks] * 4)[:12001]
    S = mpcb.sc2_sampling_lines(big.shape[0] + 1)
    sizes, st, _ = mpcb.sc2_run(big, S)
    assert np.array_equal(sizes.astype(np.uint32), oracle_sc2(big, S))
    print("ok SC2")
    sizes, st, _ = mpcb.pattern_run(big)
    want, words = oracle_pattern(big)
    assert np.array_equal(sizes.astype(np.uint32), want) and np.array_equal(st.words(), words)
    print("ok PATTERN", st.distinct_blocks, st.temporal_bytes)


if __name__ == "__main__":
    main()
