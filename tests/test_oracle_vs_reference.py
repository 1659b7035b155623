"""Oracle vs the unmodified reference classes (oracle/_ref/libmpcref.so) on random configs and data.
Skipped where the reference build is absent (it is built from /root/reference by oracle/build_ref.sh and
travels to the GPU box as a prebuilt file)."""
import numpy as np
import pytest

from helpers import SHIPPED, cfg_path, dump_config, random_blocks, random_config
from oracle.bridge import OracleMPC, RefCompressor, have_ref
from tools.gen_dump import synth

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")


@pytest.mark.parametrize("cfg", SHIPPED)
@pytest.mark.parametrize("kind", ["mixed_hashed", "smooth_f32", "sparse_i32"])
def test_shipped_configs(cfg, kind):
    d = synth(kind, 11, 5000, 1500, 1 << 16)
    r = OracleMPC(cfg_path(cfg)).run(d)
    sizes, sels = RefCompressor("VPC", cfg_path(cfg)).compress(d)
    assert np.array_equal(r.sizes, sizes) and np.array_equal(r.sels, sels)


@pytest.mark.parametrize("seed", range(12))
def test_random_configs(tmp_path, seed):
    rng = np.random.default_rng(1000 + seed)
    L = int(rng.choice([32, 64, 128]))
    cfg = random_config(rng, L=L)
    path = dump_config(cfg, str(tmp_path / "c.json"))
    d = random_blocks(rng, 300, L)
    r = OracleMPC(cfg).run(d)
    ref = RefCompressor("VPC", path)
    sizes, sels = ref.compress(d)
    assert np.array_equal(r.sizes, sizes) and np.array_equal(r.sels, sels)
    stat, fl, _ = ref.vpc_stats()
    assert np.array_equal(r.count, stat[:, 0]) and np.array_equal(r.res_lines, stat[:, 3])
    if L in (32, 64, 128):  # power-of-two L: integer-exact running means
        for k in range(r.count.size):
            n = int(r.res_lines[k])
            assert ((float(int(r.res_abs[k])) / L) / float(n) if n else 0.0) == fl[k, 1]
            assert ((float(int(r.res_sq[k])) / L) / float(n) if n else 0.0) == fl[k, 2]
