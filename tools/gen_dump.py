#!/usr/bin/env python3
"""Synthetic memory dumps for the BASELINE.json configs (numpy side).

Two families:
  * survey_mixed(n): the numpy-RNG recipe of BASELINE.md section 3 (config #1, 64 MiB mixed dump) -- kept so
    the survey's known answers (ratio 1.4392912514797536 under configs/P6.json) stay reproducible.
  * synth(kind, seed, first, n, total): counter-based generator, bit-identical to the device generator
    mpc_synth_device (cal_22-mpc_b200/csrc/mpc_synth.cu); every 128-byte block is a pure function of
    (kind, seed, global block index).

CLI:  gen_dump.py OUT.npy --kind mixed_hashed --blocks 524288 [--seed 1234]   (writes blocks+1 rows: the
reference's NPY loader drops the last row, LoaderNPY.cpp:28-32)
"""
import argparse
import sys

import numpy as np

KINDS = ["zero", "wordsame", "smooth_f32", "ramp_i32", "pointer", "random", "sparse_i32", "noisy_f32",
         "mixed_hashed", "mixed_regions"]
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth(kind, seed, first_block, n_blocks, total_blocks):
    """-> uint8 array (n_blocks, 128)"""
    kid = KINDS.index(kind) if isinstance(kind, str) else int(kind)
    b = np.arange(first_block, first_block + n_blocks, dtype=np.uint64)
    key = splitmix64(np.uint64(seed) ^ splitmix64(b))
    if kid == 8:
        cls = (key >> np.uint64(61)).astype(np.int64)
    elif kid == 9:
        cls = np.minimum((b * np.uint64(8)) // np.uint64(max(total_blocks, 1)), 7).astype(np.int64)
    else:
        cls = np.full(n_blocks, kid, dtype=np.int64)
    out = np.zeros((n_blocks, 32), dtype=np.uint32)
    k = np.arange(32, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for c in range(8):
            idx = np.nonzero(cls == c)[0]
            if idx.size == 0 or c == 0:
                continue
            ky = key[idx]
            t = splitmix64(ky[:, None] + k[None, :])  # (m, 32) uint64
            if c == 1:
                out[idx] = (ky >> np.uint64(16)).astype(np.uint32)[:, None]
            elif c == 2:
                base = np.uint32(0x3F800000) + ((b[idx] * np.uint64(2741)) & np.uint64(0x3FFFFF)).astype(np.uint32)
                d = (t & np.uint64(0x1FFF)).astype(np.uint32) - np.uint32(0x1000)
                out[idx] = base[:, None] + np.cumsum(d, axis=1, dtype=np.uint32)
            elif c == 3:
                base = (ky & np.uint64(0xFFFFF)).astype(np.uint32)
                out[idx] = base[:, None] + (np.uint32(4) * k.astype(np.uint32))[None, :]
            elif c == 4:
                v = np.uint64(0x00007F0000000000) + np.uint64(8) * (t[:, :16] & np.uint64(0x3FFFFFFF))
                out[idx] = np.ascontiguousarray(v).view(np.uint32)
            elif c == 5:
                out[idx] = t.astype(np.uint32)
            elif c == 6:
                val = ((t & np.uint64(0xFFFF)) % np.uint64(200)).astype(np.int64) - 100
                zero = ((t >> np.uint64(32)) % np.uint64(10)) < np.uint64(6)
                out[idx] = np.where(zero, 0, val).astype(np.int32).view(np.uint32)
            else:
                e = np.uint32(120) + ((t >> np.uint64(40)) % np.uint64(14)).astype(np.uint32)
                out[idx] = ((t >> np.uint64(63)).astype(np.uint32) << np.uint32(31)) | (e << np.uint32(23)) | \
                    (t & np.uint64(0x7FFFFF)).astype(np.uint32)
    return out.view(np.uint8).reshape(n_blocks, 128)


def survey_mixed(n=524289, seed=1234):
    """BASELINE.md section 3 recipe, verbatim draw order.  Returns uint8 (n, 128); the reference compresses n-1 rows."""
    L = 128
    rng = np.random.default_rng(seed)
    k = n // 8
    out = np.zeros((n, L), np.uint8)
    out[1 * k:2 * k] = np.repeat(rng.integers(0, 2**32, (k, 1), dtype=np.uint64).astype(np.uint32), 32, 1).view(np.uint8)
    x = np.cumsum(rng.normal(0, 1e-3, (k, 32)), 1).astype(np.float32) + 1.0
    out[2 * k:3 * k] = x.view(np.uint8)
    out[3 * k:4 * k] = (rng.integers(0, 1 << 20, (k, 1)) + np.arange(32)[None, :] * 4).astype(np.int32).view(np.uint8)
    p = (0x00007f0000000000 + rng.integers(0, 1 << 30, (k, 16)) * 8).astype(np.uint64)
    out[4 * k:5 * k] = p.view(np.uint8)
    out[5 * k:6 * k] = rng.integers(0, 256, (k, L), dtype=np.uint8)
    s = rng.integers(-100, 100, (k, 32)).astype(np.int32)
    s[rng.random((k, 32)) < .6] = 0
    out[6 * k:7 * k] = s.view(np.uint8)
    out[7 * k:] = rng.normal(0, 1, (n - 7 * k, 32)).astype(np.float32).view(np.uint8)
    return out


def kat_blocks():
    """The nine known-answer blocks of SURVEY.md section 8c."""
    b = np.zeros((9, 128), np.uint8)
    b[1] = np.full(32, 0xDEADBEEF, np.uint32).view(np.uint8)
    b[2] = np.arange(32, dtype=np.int32).view(np.uint8)
    b[3] = np.random.default_rng(7).integers(0, 256, 128, dtype=np.uint8)
    b[4] = np.linspace(1.0, 1.031, 32, dtype=np.float32).view(np.uint8)
    b[5] = (0x00007f1234560000 + 64 * np.arange(16, dtype=np.uint64)).view(np.uint8)
    z = np.zeros(32, np.int32)
    z[5], z[17] = -1, 3
    b[6] = z.view(np.uint8)
    r = np.zeros(32, np.int32)
    r[:16] = np.arange(16)
    b[7] = r.view(np.uint8)
    b[8] = 0xFF
    return b


GPGPUSIM_KEYS = [("kid", 1), ("mftype", 1), ("cycle", 8), ("tpc", 4), ("sid", 4), ("wid", 4), ("pc", 4), ("instcn", 4),
                 ("addr", 8), ("reqtyp", 4), ("row", 4), ("chip", 4), ("bank", 4), ("col", 4), ("reqsiz", 4), ("data", 0),
                 ("pad", 0)]


def write_gpgpusim_log(path, blocks, req_types, truncate_last=0):
    """Binary GPGPU-Sim trace as the reference's gpgpusim::LoaderGPGPU reads it (LoaderGPGPU.cpp:26-54, 82-113):
    1 byte key count (17), 17 x (6-char key, 1-byte size), then per record 62 header bytes + req_size payload bytes.
    req_types[i] is the record's request type (0 = GLOBAL_ACC_R and 4 = GLOBAL_ACC_W are the ones the driver keeps,
    main.cpp:222-224).  truncate_last > 0 cuts that many bytes off the end (an incomplete last record)."""
    import struct
    out = bytearray()
    out += bytes([len(GPGPUSIM_KEYS)])
    for name, size in GPGPUSIM_KEYS:
        out += name.encode().ljust(6, b"\0")[:6] + bytes([size])
    for i, (blk, rt) in enumerate(zip(blocks, req_types)):
        payload = bytes(blk)
        out += struct.pack("<BBQIIIIIQIIIIII", i % 3, 0 if rt < 4 else 1, 1000 + 7 * i, i % 8, i % 16, i % 48, 0x100 + 8 * i, i,
                           0x7F0000000000 + 128 * i, int(rt), i % 1024, i % 4, i % 16, i % 64, len(payload))
        out += payload
    if truncate_last:
        out = out[:-truncate_last]
    with open(path, "wb") as f:
        f.write(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--kind", default="mixed_hashed", choices=KINDS + ["survey_mixed"])
    ap.add_argument("--blocks", type=int, default=524288)
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    if a.kind == "survey_mixed":
        arr = survey_mixed(a.blocks + 1, a.seed)
    else:
        arr = np.concatenate([synth(a.kind, a.seed, 0, a.blocks, a.blocks), np.zeros((1, 128), np.uint8)])
    np.save(a.out, arr)
    print(f"wrote {a.out}: {arr.shape[0]} rows ({a.blocks} compressed by the reference)", file=sys.stderr)


if __name__ == "__main__":
    main()
