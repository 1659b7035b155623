"""Shared helpers for the tests: seeded random configs in the reference's JSON schema, and comparisons."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIPPED = ["P6", "F4", "Z1", "E5"]


def cfg_path(name):
    return os.path.join(ROOT, "configs", name + ".json")


def random_config(rng, L=128, n_pred=None, allow_root=True, table="random"):
    """A valid config with random predictors and scan tables (exercises the generic paths)."""
    n_pred = n_pred if n_pred is not None else int(rng.integers(0, 5))
    has_ws = bool(rng.integers(0, 2))
    mods = [{"name": "AllZero"}]
    if has_ws:
        mods.append({"name": "AllWordSame" if rng.integers(0, 2) else "ByteplaneAllSame"})
    for _ in range(n_pred):
        kind = int(rng.integers(0, 4))
        root = int(rng.integers(0, L)) if (allow_root and kind != 1 and rng.integers(0, 3) == 0) else 0
        pred = {"LineSize": L, "RootIndex": root}
        if kind == 0:
            pred["name"] = "OneBasePredictor"
        elif kind == 1:
            pred["name"] = "ConsecutiveBasePredictor"
        elif kind == 2:
            pred["name"] = "DiffBasePredictor"
            mode = int(rng.integers(0, 3))
            if mode == 0:
                stride = int(rng.choice([1, 2, 4, 8, 16]))
                pred["BaseIndexTable"] = [0 if i < stride else i - stride for i in range(L)]
            else:
                pred["BaseIndexTable"] = [int(v) for v in rng.integers(0, L, L)]
            pred["DiffTable"] = [0] * L if mode == 1 else [int(v) for v in rng.integers(-300, 300, L)]
        else:
            pred["name"] = "WeightBasePredictor"
            pred["BaseIndexTable"] = [int(v) for v in rng.integers(0, L, L)]
            pred["WeightTable"] = [float(v) for v in rng.choice([0.01, 0.1, 0.25, 0.3, 0.5, 1.0, 1.9, 2.0, 4.0, 100.0, 300.0], L)]
        mode = table if table != "random" else str(rng.choice(["perm", "cm", "pm", "short", "dup"]))
        nb = 8 * L
        if mode == "cm":
            cols = rng.permutation(L)
            rows_, cols_ = np.tile(np.arange(8), L), np.repeat(cols, 8)
        elif mode in ("pm", "pmr"):
            cols = rng.permutation(L)
            planes = rng.permutation(8) if mode == "pmr" else np.arange(8)
            rows_, cols_ = np.repeat(planes, L), np.tile(cols, 8)
        elif mode == "cms":  # column-major over a subset of the columns (short table)
            cols = rng.permutation(L)[: int(rng.integers(1, L))]
            rows_, cols_ = np.tile(np.arange(8), len(cols)), np.repeat(cols, 8)
        else:
            idx = rng.permutation(nb)
            if mode == "dup":
                idx = rng.integers(0, nb, nb)
            if mode == "short":
                idx = idx[: int(rng.integers(0, nb))]
            rows_, cols_ = idx // L, idx % L
        scan = {"TableSize": int(len(rows_)), "Rows": [int(v) for v in rows_], "Cols": [int(v) for v in cols_]}
        mods.append({"name": "PredComp", "submodules": {
            "ResidueModule": {"PredictorModule": pred},
            "XORModule": {"consecutiveXOR": bool(rng.integers(0, 2))},
            "ScanModule": scan,
            "FPCModule": {"num_modules": 1, "0": {"name": "UncompressedPattern", "encodingBits": 17}}}})
    ov = {"num_modules": len(mods), "lineSize": L}
    if rng.integers(0, 2):
        ov["encoding_bits"] = [int(v) for v in rng.integers(0, 12, len(mods) + 1)]
    return {"overview": ov, "modules": {str(i): m for i, m in enumerate(mods)}}


def random_blocks(rng, n, L=128):
    """Blocks that reach every branch: zeros, repeats, sparse bits, small deltas, noise."""
    out = np.zeros((n, L), np.uint8)
    for i in range(n):
        c = int(rng.integers(0, 8))
        if c == 0:
            pass
        elif c == 1:
            out[i] = np.tile(rng.integers(0, 256, 4, dtype=np.uint8), L // 4)
        elif c == 2:
            k = int(rng.integers(1, 6))
            out[i, rng.integers(0, L, k)] = (1 << rng.integers(0, 8, k)).astype(np.uint8)
        elif c == 3:
            base = rng.integers(0, 256, 4, dtype=np.uint8)
            out[i] = (np.tile(base, L // 4) + rng.integers(0, 3, L)).astype(np.uint8)
        elif c == 4:
            out[i] = np.cumsum(rng.integers(-2, 3, L)).astype(np.uint8)
        elif c == 5:
            w = (int(rng.integers(0, 1 << 20)) + np.arange(L // 4) * int(rng.integers(0, 9))).astype(np.uint32)
            out[i] = w.view(np.uint8)
        elif c == 6:
            out[i] = rng.integers(0, 256, L, dtype=np.uint8)
        else:
            out[i] = rng.integers(0, 256, L, dtype=np.uint8) & np.uint8(1 << int(rng.integers(0, 8)))
    return out


def dump_config(cfg, path):
    with open(path, "w") as f:
        json.dump(cfg, f)
    return path
