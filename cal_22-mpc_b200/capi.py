"""ctypes binding of libmpc_b200.so (include/mpc_capi.h).

Thin by design: the product is the CUDA library behind the C ABI; this module only lets
tests/ and bench.py drive it from Python.  There is no Python or CPU implementation of the
path here -- if the shared library or a CUDA device is missing, calls raise MpcError.
"""
import ctypes as C
import os

import numpy as np

MAX_LINE = 128
MAX_MODULES = 16
HIST_BINS = 8 * MAX_LINE + 32
STATS_WORDS = 2 * (MAX_MODULES + 1) + (MAX_MODULES + 1) * HIST_BINS
COMM_UID_BYTES = 128

SYN = {"zero": 0, "wordsame": 1, "smooth_f32": 2, "ramp_i32": 3, "pointer": 4, "random": 5,
       "sparse_i32": 6, "noisy_f32": 7, "mixed_hashed": 8, "mixed_regions": 9}


class ModulePod(C.Structure):
    _fields_ = [("kind", C.c_int32), ("predictor", C.c_int32), ("root", C.c_int32),
                ("consecutive_xor", C.c_int32), ("table_size", C.c_int32),
                ("base", C.c_uint8 * MAX_LINE), ("diff", C.c_uint8 * MAX_LINE), ("shift", C.c_int8 * MAX_LINE),
                ("scan_row", C.c_uint8 * (8 * MAX_LINE)), ("scan_col", C.c_uint8 * (8 * MAX_LINE))]


class ConfigPod(C.Structure):
    _fields_ = [("line_size", C.c_int32), ("num_modules", C.c_int32), ("has_wordsame", C.c_int32),
                ("first_predcomp", C.c_int32), ("enc_bits", C.c_int32 * (MAX_MODULES + 1)),
                ("modules", ModulePod * MAX_MODULES)]


class StatsPod(C.Structure):
    _fields_ = [("blocks", C.c_uint64), ("original_bits", C.c_uint64), ("compressed_bits", C.c_uint64),
                ("count", C.c_uint64 * (MAX_MODULES + 1)), ("comp_bits", C.c_uint64 * (MAX_MODULES + 1)),
                ("res_lines", C.c_uint64 * (MAX_MODULES + 1)), ("res_abs", C.c_uint64 * (MAX_MODULES + 1)),
                ("res_sq", C.c_uint64 * (MAX_MODULES + 1)),
                ("hist", (C.c_uint64 * HIST_BINS) * (MAX_MODULES + 1))]


class PatternStats(C.Structure):
    """mpc_pattern_stats (include/mpc_capi.h)"""
    _fields_ = [("blocks", C.c_uint64), ("total_bytes", C.c_uint64), ("zeros_bytes", C.c_uint64),
                ("repeated_bytes", C.c_uint64), ("temporal_bytes", C.c_uint64), ("undefined_bytes", C.c_uint64),
                ("implicit_bytes", C.c_uint64 * 6), ("explicit_bytes", C.c_uint64 * 6),
                ("symbol_counts", C.c_uint64 * 256), ("symbol_counts_nontrivial", C.c_uint64 * 256),
                ("distinct_blocks", C.c_uint64), ("temporal_path", C.c_int32)]

    def words(self):
        """the counters in the layout of the CPU oracle (orc_pattern_run)"""
        return np.array([self.zeros_bytes, self.repeated_bytes, self.temporal_bytes, self.undefined_bytes, self.total_bytes]
                        + list(self.implicit_bytes) + list(self.explicit_bytes) + list(self.symbol_counts)
                        + list(self.symbol_counts_nontrivial), dtype=np.uint64)


class Sc2Table(C.Structure):
    """mpc_sc2_table (include/mpc_capi.h)"""
    _fields_ = [("n", C.c_uint32), ("symbols", C.c_uint32 * 1024), ("lengths", C.c_uint8 * 1024)]


class VariantStats(C.Structure):
    _fields_ = [("blocks", C.c_uint64), ("original_bits", C.c_uint64), ("compressed_bits", C.c_uint64),
                ("counts", C.c_uint64 * 16)]


ALG = {"BDI": 1, "FPC": 2, "BPC": 3}


class MpcError(RuntimeError):
    pass


_LIB = None
# MPC_B200_LIB: another build of the same library (A/B runs of two kernel versions on one GPU box)
LIB_PATH = os.environ.get("MPC_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmpc_b200.so")

# every symbol include/mpc_capi.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = ["mpc_config_from_json_file", "mpc_config_from_json_text", "mpc_config_validate", "mpc_create",
           "mpc_destroy", "mpc_last_error", "mpc_global_error", "mpc_set_kernel", "mpc_kernel_name", "mpc_set_stream",
           "mpc_submit_device", "mpc_submit_host", "mpc_sync", "mpc_stats_device_ptr", "mpc_finish",
           "mpc_stats_expand", "mpc_reset", "mpc_last_timing", "mpc_synth_device", "mpc_version",
           "mpc_variant_run_device", "mpc_variant_run_host", "mpc_variant_error", "mpc_sc2_run_device", "mpc_sc2_run_host",
           "mpc_sc2_error", "mpc_cpack_run_host", "mpc_jit_compile_check", "mpc_enable_timing",
           "mpc_pattern_run_device", "mpc_pattern_run_host", "mpc_pattern_error",
           "mpc_comm_unique_id", "mpc_comm_init_rank", "mpc_comm_init_all", "mpc_attach_comm", "mpc_allreduce_stats",
           "mpc_reduced_stats", "mpc_reduced_device_ptr", "mpc_finish_allreduce", "mpc_submit_file", "mpc_prepare_host", "mpc_sc2_build_table", "mpc_sc2_apply_device"]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise MpcError(f"{LIB_PATH} is missing: build it with `make` (nvcc, sm_100a); there is no fallback path")
    l = C.CDLL(LIB_PATH)
    vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
    l.mpc_config_from_json_file.argtypes = [C.c_char_p, C.POINTER(ConfigPod), C.c_char_p, sz]
    l.mpc_config_from_json_text.argtypes = [C.c_char_p, C.POINTER(ConfigPod), C.c_char_p, sz]
    l.mpc_config_validate.argtypes = [C.POINTER(ConfigPod), C.c_char_p, sz]
    l.mpc_create.argtypes = [C.POINTER(ConfigPod), C.c_int, C.POINTER(vp)]
    l.mpc_destroy.argtypes = [vp]
    l.mpc_destroy.restype = None
    l.mpc_last_error.argtypes = [vp]
    l.mpc_last_error.restype = C.c_char_p
    l.mpc_global_error.restype = C.c_char_p
    l.mpc_set_kernel.argtypes = [vp, C.c_int]
    l.mpc_kernel_name.argtypes = [vp]
    l.mpc_set_stream.argtypes = [vp, vp]
    l.mpc_enable_timing.argtypes = [vp, C.c_int]
    l.mpc_jit_compile_check.argtypes = [C.POINTER(ConfigPod), C.c_char_p, sz, C.POINTER(sz)]
    l.mpc_kernel_name.restype = C.c_char_p
    l.mpc_submit_device.argtypes = [vp, vp, u64, vp]
    l.mpc_submit_host.argtypes = [vp, vp, u64, vp]
    l.mpc_submit_file.argtypes = [vp, C.c_int, u64, u64, vp, C.c_int]
    l.mpc_prepare_host.argtypes = [vp]
    l.mpc_sync.argtypes = [vp]
    l.mpc_stats_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    l.mpc_finish.argtypes = [vp, C.POINTER(StatsPod)]
    l.mpc_stats_expand.argtypes = [C.POINTER(ConfigPod), vp, sz, C.POINTER(StatsPod)]
    l.mpc_reset.argtypes = [vp]
    l.mpc_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    l.mpc_synth_device.argtypes = [vp, vp, u64, u64, u64, C.c_int, u64]
    l.mpc_version.restype = C.c_char_p
    l.mpc_variant_run_device.argtypes = [C.c_int, C.c_int, vp, u64, C.c_uint32, vp, C.POINTER(VariantStats), C.POINTER(C.c_float)]
    l.mpc_variant_run_host.argtypes = [C.c_int, C.c_int, vp, u64, C.c_uint32, vp, C.POINTER(VariantStats), C.POINTER(C.c_float)]
    l.mpc_variant_error.restype = C.c_char_p
    l.mpc_sc2_run_device.argtypes = [C.c_int, vp, u64, C.c_uint32, u64, vp, C.POINTER(VariantStats), C.POINTER(C.c_float)]
    l.mpc_sc2_run_host.argtypes = [C.c_int, vp, u64, C.c_uint32, u64, vp, C.POINTER(VariantStats), C.POINTER(C.c_float)]
    l.mpc_sc2_error.restype = C.c_char_p
    l.mpc_sc2_build_table.argtypes = [C.c_int, vp, u64, C.c_uint32, C.POINTER(Sc2Table)]
    l.mpc_sc2_apply_device.argtypes = [C.c_int, vp, u64, u64, u64, C.c_uint32, C.POINTER(Sc2Table), vp, C.POINTER(VariantStats), C.POINTER(C.c_float)]
    l.mpc_cpack_run_host.argtypes = [vp, u64, C.c_uint32, vp, C.POINTER(VariantStats)]
    l.mpc_pattern_run_device.argtypes = [C.c_int, vp, u64, C.c_uint32, u64, vp, C.POINTER(PatternStats), C.POINTER(C.c_float)]
    l.mpc_pattern_run_host.argtypes = [C.c_int, vp, u64, C.c_uint32, u64, vp, C.POINTER(PatternStats), C.POINTER(C.c_float)]
    l.mpc_pattern_error.restype = C.c_char_p
    l.mpc_comm_unique_id.argtypes = [vp, sz]
    l.mpc_comm_init_rank.argtypes = [vp, vp, sz, C.c_int, C.c_int]
    l.mpc_comm_init_all.argtypes = [C.POINTER(vp), C.c_int]
    l.mpc_attach_comm.argtypes = [vp, vp]
    l.mpc_allreduce_stats.argtypes = [C.POINTER(vp), C.c_int]
    l.mpc_reduced_stats.argtypes = [vp, C.POINTER(StatsPod)]
    l.mpc_reduced_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    l.mpc_finish_allreduce.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(StatsPod)]
    _LIB = l
    return l


def load_config(path=None, text=None):
    """JSON (reference format, VPC.cpp:72-330) -> ConfigPod; raises MpcError with the library's message."""
    pod = ConfigPod()
    err = C.create_string_buffer(1024)
    if path is not None:
        rc = lib().mpc_config_from_json_file(os.fsencode(path), C.byref(pod), err, 1024)
    else:
        rc = lib().mpc_config_from_json_text(text.encode(), C.byref(pod), err, 1024)
    if rc != 0:
        raise MpcError(f"config rejected ({rc}): {err.value.decode(errors='replace')}")
    return pod


def jit_compile_check(config):
    """Config compiler + NVRTC on the CPU: -> (return code, cubin bytes, log)."""
    pod = config if isinstance(config, ConfigPod) else load_config(path=config)
    log = C.create_string_buffer(1 << 16)
    n = C.c_size_t()
    rc = lib().mpc_jit_compile_check(C.byref(pod), log, len(log), C.byref(n))
    return rc, n.value, log.value.decode(errors="replace")


class Stats:
    """Expanded statistics; field names follow comp::VPCResult / comp::CompResult."""

    def __init__(self, pod, num_modules, line_size):
        k = num_modules + 1
        self.blocks = int(pod.blocks)
        self.OriginalSize = int(pod.original_bits)
        self.CompressedSize = int(pod.compressed_bits)
        self.count = np.array(pod.count[:k], dtype=np.uint64)
        self.comp_bits = np.array(pod.comp_bits[:k], dtype=np.uint64)
        self.res_lines = np.array(pod.res_lines[:k], dtype=np.uint64)
        self.res_abs = np.array(pod.res_abs[:k], dtype=np.uint64)
        self.res_sq = np.array(pod.res_sq[:k], dtype=np.uint64)
        self.hist = np.ctypeslib.as_array(pod.hist)[:k].copy()
        self.line_size = line_size

    @property
    def CompRatio(self):  # CompResult.h:34
        return float(self.OriginalSize) / float(self.CompressedSize) if self.CompressedSize else float("nan")

    def mae(self, k):  # VPC.h:62-76 with the exact reformulation of SURVEY.md section 7
        n = int(self.res_lines[k])
        return (float(int(self.res_abs[k])) / self.line_size) / float(n) if n else 0.0

    def mse(self, k):
        n = int(self.res_lines[k])
        return (float(int(self.res_sq[k])) / self.line_size) / float(n) if n else 0.0


class Mpc:
    """One context = one GPU (mpc_create .. mpc_destroy)."""

    def __init__(self, config, device=0):
        self.cfg = config if isinstance(config, ConfigPod) else load_config(path=config)
        h = C.c_void_p()
        rc = lib().mpc_create(C.byref(self.cfg), device, C.byref(h))
        if rc != 0:
            raise MpcError(f"mpc_create failed ({rc}): {lib().mpc_global_error().decode()}")
        self.h = h
        self.line_size = self.cfg.line_size

    def _check(self, rc):
        if rc != 0:
            raise MpcError(f"libmpc_b200 error {rc}: {lib().mpc_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            lib().mpc_destroy(self.h)
            self.h = None

    __del__ = close

    def set_kernel(self, which):
        self._check(lib().mpc_set_kernel(self.h, which))

    def set_stream(self, cuda_stream):
        self._check(lib().mpc_set_stream(self.h, cuda_stream))

    def enable_timing(self, enabled):
        self._check(lib().mpc_enable_timing(self.h, 1 if enabled else 0))

    def kernel_name(self):
        return lib().mpc_kernel_name(self.h).decode()

    def submit_device(self, d_lines_ptr, n_blocks, d_packed_ptr=None):
        self._check(lib().mpc_submit_device(self.h, d_lines_ptr, n_blocks, d_packed_ptr))

    def submit_host(self, lines, packed=None):
        lines = np.ascontiguousarray(lines, dtype=np.uint8)
        n = lines.size // self.line_size
        pp = packed.ctypes.data if packed is not None else None
        self._keep = (lines, packed)
        self._check(lib().mpc_submit_host(self.h, lines.ctypes.data, n, pp))

    def submit_host_ptr(self, ptr, n_blocks, packed_ptr=None):
        self._check(lib().mpc_submit_host(self.h, ptr, n_blocks, packed_ptr))

    def submit_file(self, fd, file_offset, n_blocks, packed=None, direct_io=False):
        pp = packed.ctypes.data if packed is not None else None
        self._keep = (packed,)
        self._check(lib().mpc_submit_file(self.h, fd, file_offset, n_blocks, pp, 1 if direct_io else 0))

    def sync(self):
        self._check(lib().mpc_sync(self.h))

    def reset(self):
        self._check(lib().mpc_reset(self.h))

    def finish(self):
        pod = StatsPod()
        self._check(lib().mpc_finish(self.h, C.byref(pod)))
        return Stats(pod, self.cfg.num_modules, self.line_size)

    def stats_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(lib().mpc_stats_device_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def expand(self, words):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        pod = StatsPod()
        self._check(lib().mpc_stats_expand(C.byref(self.cfg), words.ctypes.data, words.size, C.byref(pod)))
        return Stats(pod, self.cfg.num_modules, self.line_size)

    # ---- multi-GPU: the statistics all-reduce behind the ABI (one context per rank) ----
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(COMM_UID_BYTES)
        rc = lib().mpc_comm_unique_id(buf, COMM_UID_BYTES)
        if rc != 0:
            raise MpcError(f"mpc_comm_unique_id failed ({rc}): {lib().mpc_global_error().decode()}")
        return buf.raw

    def comm_init_rank(self, uid, nranks, rank):
        self._check(lib().mpc_comm_init_rank(self.h, uid, len(uid), nranks, rank))

    def allreduce_stats(self):
        """asynchronous on the context's stream: all-reduce a copy of the statistics vector"""
        arr = (C.c_void_p * 1)(self.h)
        self._check(lib().mpc_allreduce_stats(arr, 1))

    def reduced_stats(self):
        pod = StatsPod()
        self._check(lib().mpc_reduced_stats(self.h, C.byref(pod)))
        return Stats(pod, self.cfg.num_modules, self.line_size)

    def finish_allreduce(self):
        pod = StatsPod()
        arr = (C.c_void_p * 1)(self.h)
        self._check(lib().mpc_finish_allreduce(arr, 1, C.byref(pod)))
        return Stats(pod, self.cfg.num_modules, self.line_size)

    def last_timing(self):
        ms, n = C.c_float(), C.c_int()
        self._check(lib().mpc_last_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def synth_device(self, d_ptr, first_block, n_blocks, total_blocks, kind, seed):
        k = SYN[kind] if isinstance(kind, str) else kind
        self._check(lib().mpc_synth_device(self.h, d_ptr, first_block, n_blocks, total_blocks, k, seed))

    def compress(self, lines):
        """Convenience: host array of blocks -> (sizes, selected) per block + Stats (fresh statistics)."""
        lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, self.line_size)
        packed = np.zeros(lines.shape[0], dtype=np.uint16)
        self.reset()
        self.submit_host(lines, packed)
        st = self.finish()
        return (packed & 0x7FF).astype(np.uint32), (packed >> 11).astype(np.int32) - 1, st


def variant_run(alg, lines=None, device_ptr=None, n_blocks=None, device=0, want_sizes=True, line_size=128):
    """BDI / FPC / BPC over host blocks (numpy) or a device pointer -> (sizes or None, VariantStats, kernel_ms)."""
    st, ms = VariantStats(), C.c_float()
    aid = ALG[alg] if isinstance(alg, str) else alg
    if device_ptr is None:
        lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, line_size)
        n = lines.shape[0]
        sizes = np.zeros(n, dtype=np.uint16) if want_sizes else None
        rc = lib().mpc_variant_run_host(aid, device, lines.ctypes.data, n, line_size,
                                        sizes.ctypes.data if want_sizes else None, C.byref(st), C.byref(ms))
    else:
        sizes = None
        rc = lib().mpc_variant_run_device(aid, device, device_ptr, n_blocks, line_size, None, C.byref(st), C.byref(ms))
    if rc != 0:
        raise MpcError(f"variant {alg} failed ({rc}): {lib().mpc_variant_error().decode()}")
    return sizes, st, ms.value


def sc2_sampling_lines(loader_rows):
    """main.cpp:108-114: sampling count from the loader's row count."""
    return max(10000, min(loader_rows // 100, 1000000))


def sc2_run(lines, sampling_lines, device=0, line_size=128):
    lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, line_size)
    n = lines.shape[0]
    sizes = np.zeros(n, dtype=np.uint16)
    st, ms = VariantStats(), C.c_float()
    rc = lib().mpc_sc2_run_host(device, lines.ctypes.data, n, line_size, sampling_lines, sizes.ctypes.data, C.byref(st), C.byref(ms))
    if rc != 0:
        raise MpcError(f"SC2 failed ({rc}): {lib().mpc_sc2_error().decode()}")
    return sizes, st, ms.value


def cpack_run(lines, line_size=128):
    lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, line_size)
    n = lines.shape[0]
    sizes = np.zeros(n, dtype=np.uint16)
    st = VariantStats()
    rc = lib().mpc_cpack_run_host(lines.ctypes.data, n, line_size, sizes.ctypes.data, C.byref(st))
    if rc != 0:
        raise MpcError(f"CPACK failed ({rc})")
    return sizes, st


def pattern_run(lines=None, device_ptr=None, n_blocks=None, cache_blocks=0, device=0, want_sizes=True, line_size=128):
    """PATTERN analysis over host blocks (numpy) or a device pointer -> (sizes or None, PatternStats, device_ms)."""
    st, ms = PatternStats(), C.c_float()
    if device_ptr is None:
        lines = np.ascontiguousarray(lines, dtype=np.uint8).reshape(-1, line_size)
        n = lines.shape[0]
        sizes = np.zeros(n, dtype=np.uint16) if want_sizes else None
        rc = lib().mpc_pattern_run_host(device, lines.ctypes.data, n, line_size, cache_blocks,
                                        sizes.ctypes.data if want_sizes else None, C.byref(st), C.byref(ms))
    else:
        sizes = None
        rc = lib().mpc_pattern_run_device(device, device_ptr, n_blocks, line_size, cache_blocks, None, C.byref(st), C.byref(ms))
    if rc != 0:
        raise MpcError(f"PATTERN failed ({rc}): {lib().mpc_pattern_error().decode()}")
    return sizes, st, ms.value


def unpack(packed):
    packed = np.asarray(packed, dtype=np.uint16)
    return (packed & 0x7FF).astype(np.uint32), (packed >> 11).astype(np.int32) - 1
