"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the CPU oracle (oracle/libmpc_oracle.so, the plain-C
restatement) and for the unmodified reference build (oracle/_ref/libmpcref.so, see build_ref.sh).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
The oracle parses the reference-format JSON with Python's json module, independently of the product's
own parser (cal_22-mpc_b200/csrc/mpc_config.cpp), so that the parser is itself under test.
"""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmpc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmpcref.so")
REF_BIN = os.path.join(HERE, "_ref", "compressor")

PRED = {"OneBasePredictor": 0, "ConsecutiveBasePredictor": 1, "DiffBasePredictor": 2, "WeightBasePredictor": 3}

_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        l = C.CDLL(ORACLE_SO)
        l.orc_new.restype = C.c_void_p
        l.orc_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.orc_default_enc_bits.argtypes = [C.c_int]
        l.orc_add_predcomp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_void_p]
        l.orc_add_predcomp.restype = None
        l.orc_free.argtypes = [C.c_void_p]
        l.orc_free.restype = None
        l.orc_run.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_uint, C.c_int]
        l.orc_run.restype = None
        l.orc_variant_run.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p, C.c_void_p]
        l.orc_variant_run.restype = None
        _oracle = l
    return _oracle


def oracle_cpack(lines, L=128):
    lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, L)
    sizes = np.zeros(lines.shape[0], dtype=np.uint32)
    counts = np.zeros(8, dtype=np.uint64)
    l = oracle_lib()
    l.orc_cpack_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p, C.c_void_p]
    l.orc_cpack_run.restype = None
    l.orc_cpack_run(lines.ctypes.data, lines.shape[0], L, sizes.ctypes.data, counts.ctypes.data)
    return sizes, counts


def oracle_sc2(lines, sampling, L=128):
    lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, L)
    sizes = np.zeros(lines.shape[0], dtype=np.uint32)
    l = oracle_lib()
    l.orc_sc2_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64, C.c_void_p]
    l.orc_sc2_run.restype = None
    l.orc_sc2_run(lines.ctypes.data, lines.shape[0], L, sampling, sizes.ctypes.data)
    return sizes


PATTERN_WORDS = 529  # layout: see orc_pattern_run in variants_oracle.c


def oracle_pattern(lines, L=128, capacity=(1 << 24) - 1):
    """CPU oracle of the PATTERN analysis: -> (sizes uint32 [n], stats uint64 [PATTERN_WORDS])"""
    lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, L)
    sizes = np.zeros(lines.shape[0], dtype=np.uint32)
    st = np.zeros(PATTERN_WORDS, dtype=np.uint64)
    l = oracle_lib()
    l.orc_pattern_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_uint64, C.c_void_p, C.c_void_p]
    l.orc_pattern_run.restype = None
    l.orc_pattern_run(lines.ctypes.data, lines.shape[0], L, capacity, sizes.ctypes.data, st.ctypes.data)
    return sizes, st


VARIANT_ID = {"BDI": 1, "FPC": 2, "BPC": 3}


def oracle_variant(alg, lines, L=128):
    """CPU oracle of BDI / FPC / BPC: -> (sizes uint32 [n], counts uint64 [16])"""
    lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, L)
    sizes = np.zeros(lines.shape[0], dtype=np.uint32)
    counts = np.zeros(16, dtype=np.uint64)
    oracle_lib().orc_variant_run(VARIANT_ID[alg], lines.ctypes.data, lines.shape[0], L, sizes.ctypes.data, counts.ctypes.data)
    return sizes, counts


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        l = C.CDLL(REF_SO)
        l.ref_create.restype = C.c_void_p
        l.ref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_ulonglong]
        l.ref_line_size.argtypes = [C.c_void_p]
        l.ref_num_modules.argtypes = [C.c_void_p]
        l.ref_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_uint, C.c_void_p, C.c_void_p]
        l.ref_compress.restype = None
        l.ref_totals.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.ref_totals.restype = None
        l.ref_vpc_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
        l.ref_vpc_stats.restype = None
        l.ref_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        l.ref_vpc_print.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        l.ref_vpc_print.restype = None
        l.ref_pattern_stats.argtypes = [C.c_void_p, C.c_void_p]
        l.ref_pattern_stats.restype = None
        l.ref_pattern_print.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        l.ref_pattern_print.restype = None
        _ref = l
    return _ref


class OracleResult:
    pass


class OracleMPC:
    """CPU oracle configured from a reference-format config (dict or path)."""

    def __init__(self, config):
        if not isinstance(config, dict):
            with open(config) as f:
                config = json.load(f)
        self.config = config
        ov = config["overview"]
        n, L = int(ov["num_modules"]), int(ov["lineSize"])
        self.n, self.L = n, L
        mods = [config["modules"][str(i)] for i in range(n)]
        names = [m["name"] for m in mods]
        has_ws = 0
        for nm in names:  # last AllZero/AllWordSame parsed decides, VPC.cpp:312,318
            if nm == "AllZero":
                has_ws = 0
            elif nm in ("AllWordSame", "ByteplaneAllSame"):
                has_ws = 1
        self.has_wordsame = has_ws
        self.first_predcomp = 2 if has_ws else 1
        l = oracle_lib()
        if ov.get("encoding_bits") is None:
            enc = [l.orc_default_enc_bits(n)] * (n + 1)
        else:
            enc = [int(v) for v in ov["encoding_bits"][: n + 1]]
        self.enc = enc
        encarr = np.array(enc, dtype=np.int32)
        self.h = l.orc_new(L, n, has_ws, encarr.ctypes.data)
        for m in mods[self.first_predcomp:]:
            sub = m["submodules"]
            ps = sub["ResidueModule"]["PredictorModule"]
            pred = PRED[ps["name"]]
            base = np.array(ps.get("BaseIndexTable", [0] * L), dtype=np.int32)
            diff = np.array(ps.get("DiffTable", [0] * L), dtype=np.int32)
            weight = np.array(ps.get("WeightTable", [1.0] * L), dtype=np.float32)
            sc = sub["ScanModule"]
            T = int(sc["TableSize"])
            rows = np.array(sc["Rows"][:T], dtype=np.int32)
            cols = np.array(sc["Cols"][:T], dtype=np.int32)
            l.orc_add_predcomp(self.h, pred, int(ps["RootIndex"]), int(bool(sub["XORModule"]["consecutiveXOR"])),
                               base.ctypes.data, diff.ctypes.data,
                               weight.ctypes.data if pred == 3 else None, T, rows.ctypes.data, cols.ctypes.data)

    def __del__(self):
        if getattr(self, "h", None):
            oracle_lib().orc_free(self.h)
            self.h = None

    def run(self, lines, threads=0, hist_bins=None):
        lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, self.L)
        nb = lines.shape[0]
        K = self.n + 1
        hb = hist_bins or (8 * self.L + 32)
        res = OracleResult()
        res.sizes = np.zeros(nb, dtype=np.uint32)
        res.sels = np.zeros(nb, dtype=np.int32)
        st = np.zeros(3 + 5 * K, dtype=np.uint64)
        res.hist = np.zeros((K, hb), dtype=np.uint64)
        if threads <= 0:
            threads = os.cpu_count() or 1
        oracle_lib().orc_run(self.h, lines.ctypes.data, nb, res.sizes.ctypes.data, res.sels.ctypes.data,
                             st.ctypes.data, res.hist.ctypes.data, hb, threads)
        res.blocks, res.OriginalSize, res.CompressedSize = int(st[0]), int(st[1]), int(st[2])
        per = st[3:].reshape(K, 5)
        res.count, res.comp_bits, res.res_lines, res.res_abs, res.res_sq = (per[:, i].copy() for i in range(5))
        return res


class RefCompressor:
    """The unmodified reference classes behind oracle/ref_harness.cpp (single-threaded, as the reference is)."""

    def __init__(self, alg, config_path=None, line_size=128, sampling=10000):
        l = ref_lib()
        self.alg = alg
        self.h = l.ref_create(alg.encode(), os.fsencode(config_path) if config_path else b"", line_size, sampling)
        if not self.h:
            raise RuntimeError("ref_create failed for " + alg)
        self.L = l.ref_line_size(self.h)
        self.n = l.ref_num_modules(self.h)

    def compress(self, lines, want_sels=True):
        lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, self.L)
        nb = lines.shape[0]
        sizes = np.zeros(nb, dtype=np.uint32)
        sels = np.zeros(nb, dtype=np.int32)
        ref_lib().ref_compress(self.h, lines.ctypes.data, nb, self.L, sizes.ctypes.data,
                               sels.ctypes.data if (want_sels and self.alg == "VPC") else None)
        return sizes, sels

    def totals(self):
        t = np.zeros(2, dtype=np.uint64)
        r = C.c_double()
        ref_lib().ref_totals(self.h, t.ctypes.data, C.byref(r))
        return int(t[0]), int(t[1]), r.value

    def vpc_stats(self, hist_bins=1056):
        K = self.n + 1
        stat = np.zeros((K, 4), dtype=np.uint64)
        fl = np.zeros((K, 3), dtype=np.float64)
        hist = np.zeros((K, hist_bins), dtype=np.uint64)
        ref_lib().ref_vpc_stats(self.h, stat.ctypes.data, fl.ctypes.data, hist.ctypes.data, hist_bins)
        return stat, fl, hist

    def counts(self, cap=16):
        out = np.zeros(cap, dtype=np.uint64)
        n = ref_lib().ref_counts(self.h, out.ctypes.data, cap)
        return out[:n]

    def pattern_stats(self):
        out = np.zeros(PATTERN_WORDS, dtype=np.uint64)
        ref_lib().ref_pattern_stats(self.h, out.ctypes.data)
        return out

    def pattern_print(self, workload, path):
        ref_lib().ref_pattern_print(self.h, workload.encode(), os.fsencode(path))

    def vpc_print(self, workload, path, detail_path):
        ref_lib().ref_vpc_print(self.h, workload.encode(), os.fsencode(path), os.fsencode(detail_path))
