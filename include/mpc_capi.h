/*
 * mpc_capi.h -- C ABI of libmpc_b200.so, the B200 (sm_100a) implementation of the per-block
 * compression loop of MPC (scalable-arch/CAL_22-MPC, where MPC is the `VPC` compressor).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  The host
 * classes in cal_22-mpc_b200/host/ (comp::Compressor / trace::Loader mirrors) and the Python
 * ctypes binding sit above it.  Each entry point names the reference interface it replaces
 * (paths relative to the reference's src/).
 *
 * All functions return 0 on success and a negative MPC_E_* code on failure; the message is
 * available from mpc_last_error() (per context) or mpc_global_error() (when no context exists).
 * Nothing in the library aborts or exits.  There is NO CPU fallback: without a CUDA device
 * mpc_create() fails with MPC_E_CUDA.
 */
#ifndef MPC_CAPI_H_
#define MPC_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPC_MAX_LINE 128      /* bytes per block (cache line); supported: 32, 64, 128 */
#define MPC_MAX_MODULES 16    /* overview.num_modules upper bound (clusters = modules + 1) */
#define MPC_MAX_ENC_BITS 31   /* encoding bits per cluster must lie in [0, 31] */
#define MPC_HIST_BINS (8 * MPC_MAX_LINE + 32) /* compressed size of a block is < MPC_HIST_BINS */

enum { MPC_OK = 0, MPC_E_ARG = -1, MPC_E_CONFIG = -2, MPC_E_CUDA = -3, MPC_E_IO = -4, MPC_E_STATE = -5 };

/* module kinds, VPC.cpp:126-325 */
enum { MPC_MOD_ALLZERO = 0, MPC_MOD_ALLWORDSAME = 1, MPC_MOD_PREDCOMP = 2 };
/* predictors, VPCmodules/PredictorModule.cpp */
enum { MPC_PRED_ONE = 0, MPC_PRED_CONSEC = 1, MPC_PRED_DIFF = 2, MPC_PRED_WEIGHT = 3 };

/* One PredComp module, flattened from the JSON the reference parses in VPC.cpp:131-306. */
typedef struct {
  int32_t kind;            /* MPC_MOD_* */
  int32_t predictor;       /* MPC_PRED_* (PredComp only) */
  int32_t root;            /* RootIndex */
  int32_t consecutive_xor; /* XORModule.consecutiveXOR */
  int32_t table_size;      /* ScanModule.TableSize, <= 8 * line_size */
  uint8_t base[MPC_MAX_LINE];      /* BaseIndexTable (Diff/Weight) */
  uint8_t diff[MPC_MAX_LINE];      /* (uint8_t)DiffTable[i]            (Diff)   */
  int8_t shift[MPC_MAX_LINE];      /* (int)log2f(WeightTable[i])       (Weight) */
  uint8_t scan_row[8 * MPC_MAX_LINE]; /* ScanModule.Rows (bit plane 0..7, 0 = MSB) */
  uint8_t scan_col[8 * MPC_MAX_LINE]; /* ScanModule.Cols (byte column) */
} mpc_module_pod;

/* Whole config; replaces the object graph VPC::parseConfig builds (VPC.cpp:72-330). */
typedef struct {
  int32_t line_size;      /* overview.lineSize */
  int32_t num_modules;    /* overview.num_modules */
  int32_t has_wordsame;   /* module 1 is AllWordSame / ByteplaneAllSame */
  int32_t first_predcomp; /* 1 or 2 */
  int32_t enc_bits[MPC_MAX_MODULES + 1]; /* index = cluster + 1; cluster -1 = uncompressed */
  mpc_module_pod modules[MPC_MAX_MODULES];
} mpc_config_pod;

/* Statistics of everything submitted since create/reset.  Replaces comp::VPCResult
 * (VPC.h:34-238) + comp::CompResult (CompResult.h:24-86).  Cluster index k = selected + 1. */
typedef struct {
  uint64_t blocks;
  uint64_t original_bits;   /* CompResult::OriginalSize   */
  uint64_t compressed_bits; /* CompResult::CompressedSize */
  uint64_t count[MPC_MAX_MODULES + 1];      /* ClusterStat::count           */
  uint64_t comp_bits[MPC_MAX_MODULES + 1];  /* ClusterStat::compressedSize  */
  uint64_t res_lines[MPC_MAX_MODULES + 1];  /* VPCResult::m_NumLines        */
  uint64_t res_abs[MPC_MAX_MODULES + 1];    /* sum over lines of sum_i |r_i|  (MAE numerator) */
  uint64_t res_sq[MPC_MAX_MODULES + 1];     /* sum over lines of sum_i r_i^2  (MSE numerator) */
  uint64_t hist[MPC_MAX_MODULES + 1][MPC_HIST_BINS]; /* ClusterStat::compSizeHistogram */
} mpc_stats_pod;

/* number of uint64 words in the device-side statistics vector (the buffer an allreduce sums) */
#define MPC_STATS_WORDS ((size_t)(2 * (MPC_MAX_MODULES + 1) + (MPC_MAX_MODULES + 1) * MPC_HIST_BINS))

typedef struct mpc_ctx mpc_ctx;

/* ---- config ---------------------------------------------------------------------------- */

/* Parse + validate a reference-format JSON config (VPC::parseConfig, VPC.cpp:72-330).
 * Configs the reference would run with undefined behaviour are rejected with MPC_E_CONFIG. */
int mpc_config_from_json_file(const char* path, mpc_config_pod* out, char* err, size_t err_len);
int mpc_config_from_json_text(const char* text, mpc_config_pod* out, char* err, size_t err_len);
int mpc_config_validate(const mpc_config_pod* cfg, char* err, size_t err_len);

/* ---- context --------------------------------------------------------------------------- */

/* Replaces `new comp::VPC(configPath)` (main.cpp:88-91).  One context drives one GPU. */
int mpc_create(const mpc_config_pod* cfg, int device, mpc_ctx** out);
void mpc_destroy(mpc_ctx* ctx);
const char* mpc_last_error(const mpc_ctx* ctx);
const char* mpc_global_error(void);

/* kernel selection: 0 = auto (specialised kernel when one exists for the config, else generic),
 * 1 = generic warp-per-block kernel, 2 = specialised thread-per-block kernel (error if none).
 * A specialised kernel exists for the configs compiled into the library and, for any other config with 128-byte
 * lines and column- or plane-major scan tables, is built at mpc_create time with NVRTC (a few seconds; MPC_JIT=0
 * disables it, MPC_JIT_CACHE_DIR=<dir> keeps the cubins on disk). */
int mpc_set_kernel(mpc_ctx* ctx, int which);
/* name of the kernel the next submit will launch ("generic_warp", "spec_thread:<cfg>", "spec_thread:jit") */
const char* mpc_kernel_name(const mpc_ctx* ctx);

/* Run the config compiler + NVRTC for cfg without touching a device (build check / tooling): MPC_OK and the size of
 * the sm_100a cubin, MPC_E_CONFIG when the config is not eligible, MPC_E_STATE + compiler log on a build error. */
int mpc_jit_compile_check(const mpc_config_pod* cfg, char* log, size_t log_len, size_t* cubin_bytes);

/* Run mpc_submit_device / mpc_synth_device / statistics resets on the caller's CUDA stream (a cudaStream_t
 * passed as void*; NULL restores the context's own stream), so that the caller's events and collectives
 * order against the kernels. */
int mpc_set_stream(mpc_ctx* ctx, void* cuda_stream);

/* ---- the hot path: replaces the compressLines loop (main.cpp:229-244) -------------------- */

/* Blocks already resident in HBM.  d_lines: device pointer, 16-byte aligned, n_blocks*line_size
 * bytes.  d_packed (optional, device): one uint16 per block = size_bits | (selected+1) << 11.
 * Asynchronous on the context's stream; statistics accumulate on the device. */
int mpc_submit_device(mpc_ctx* ctx, const uint8_t* d_lines, uint64_t n_blocks, uint16_t* d_packed);

/* Blocks in host memory (pageable or pinned).  Chunked through a ring of three pinned
 * staging slots: cudaMemcpyAsync H2D on rotating streams overlapped with the kernel.  h_packed (optional,
 * host) receives the per-block results.  Returns after the last chunk has been issued;
 * mpc_finish() / mpc_sync() wait for completion. */
int mpc_submit_host(mpc_ctx* ctx, const uint8_t* h_lines, uint64_t n_blocks, uint16_t* h_packed);

/* Blocks in a FILE: n_blocks * line_size bytes starting at file_offset of the open descriptor fd (the data part of an .npy
 * dump, a raw dump ...).  Replaces LoaderNPY's "read the whole file into a vector, then copy one line per call"
 * (LoaderNPY.cpp:14-34, 48-54) for dumps of any size, including dumps larger than host memory: the file is read with
 * pread() on several threads STRAIGHT into the context's pinned staging ring (three 32 MiB slots: while one slot is
 * being filled, the previous one is on its way over PCIe and the one before is being compressed), so the data is touched
 * once on the host and never mapped.  direct_io != 0: fd was opened with O_DIRECT (cold dumps that should not go through
 * the page cache); offsets and lengths are aligned internally.  Returns after the last chunk has been issued. */
int mpc_submit_file(mpc_ctx* ctx, int fd, uint64_t file_offset, uint64_t n_blocks, uint16_t* h_packed, int direct_io);

/* Allocate the staging ring of mpc_submit_host / mpc_submit_file now (pinning host memory takes tens of milliseconds) instead
 * of inside the first submit. */
int mpc_prepare_host(mpc_ctx* ctx);

/* Wait for all submitted work. */
int mpc_sync(mpc_ctx* ctx);

/* Device statistics vector (MPC_STATS_WORDS uint64, layout private to the library but linear:
 * element-wise sums of vectors from several GPUs are valid).  This is the buffer a multi-GPU
 * caller all-reduces (ncclAllReduce, ncclUint64/ncclInt64, ncclSum) before mpc_finish(). */
int mpc_stats_device_ptr(mpc_ctx* ctx, uint64_t** d_stats, size_t* n_words);

/* Sync, copy the statistics vector back and expand it.  Replaces compressor->GetResult()
 * (Compressor.h:29, main.cpp:246). */
int mpc_finish(mpc_ctx* ctx, mpc_stats_pod* out);

/* Expand a host copy of a statistics vector (e.g. after an allreduce done elsewhere). */
int mpc_stats_expand(const mpc_config_pod* cfg, const uint64_t* words, size_t n_words, mpc_stats_pod* out);

/* ---- multi-GPU: blocks are independent, so every GPU compresses a contiguous shard of the dump
 * and the job's ONE exchange is an all-reduce (sum, uint64) of the statistics vector over NVLink.  The reference has no
 * counterpart (single process, single thread; VPCResult::Update, VPC.h:49-76, is the state that is summed).
 * One context per GPU.  Multi-process jobs (one rank per GPU): rank 0 calls mpc_comm_unique_id, ships the bytes to the
 * other ranks by any means, every rank calls mpc_comm_init_rank.  Single-process jobs: mpc_comm_init_all over the
 * process's contexts.  A caller that already has an ncclComm_t can lend it with mpc_attach_comm (not destroyed by
 * mpc_destroy).  NCCL is loaded with dlopen on first use; every NCCL return code is checked and reported through
 * mpc_last_error(). */
#define MPC_COMM_UID_BYTES 128
int mpc_comm_unique_id(void* uid, size_t uid_len);
int mpc_comm_init_rank(mpc_ctx* ctx, const void* uid, size_t uid_len, int nranks, int rank);
int mpc_comm_init_all(mpc_ctx** ctxs, int n);
int mpc_attach_comm(mpc_ctx* ctx, void* nccl_comm);
/* Asynchronous, on each context's stream: copy the statistics vector aside and all-reduce the copy over the
 * communicator (ncclAllReduce, ncclUint64, ncclSum; one NCCL group over the n local contexts).  The local vector keeps
 * accumulating, so the call can be repeated.  Contexts without a communicator (a one-GPU job): the copy is the result. */
int mpc_allreduce_stats(mpc_ctx** ctxs, int n);
/* Wait for the context's stream, copy the all-reduced vector back and expand it. */
int mpc_reduced_stats(mpc_ctx* ctx, mpc_stats_pod* out);
/* the all-reduced vector itself (MPC_STATS_WORDS uint64, device memory) */
int mpc_reduced_device_ptr(mpc_ctx* ctx, uint64_t** d_reduced, size_t* n_words);
/* mpc_sync on every context + mpc_allreduce_stats + mpc_reduced_stats(ctxs[0]): the multi-GPU form of mpc_finish(). */
int mpc_finish_allreduce(mpc_ctx** ctxs, int n, mpc_stats_pod* out);

/* Zero the device statistics (new file / new run). */
int mpc_reset(mpc_ctx* ctx);

/* Device time (CUDA events on the context's stream) of the kernels launched by the most recent
 * mpc_submit_device / mpc_submit_host call, and the number of kernel launches it made. */
int mpc_last_timing(mpc_ctx* ctx, float* kernel_ms, int* launches);

/* mpc_submit_device brackets its launch with two CUDA events for mpc_last_timing; a caller that times a train of
 * back-to-back launches itself can switch them off (enabled = 0) to keep the stream free of extra commands. */
int mpc_enable_timing(mpc_ctx* ctx, int enabled);

/* ---- synthetic dumps (BASELINE.json configs; generator shared with tools/gen_dump.py) ---- */

enum {
  MPC_SYN_ZERO = 0, MPC_SYN_WORDSAME = 1, MPC_SYN_SMOOTH_F32 = 2, MPC_SYN_RAMP_I32 = 3,
  MPC_SYN_POINTER = 4, MPC_SYN_RANDOM = 5, MPC_SYN_SPARSE_I32 = 6, MPC_SYN_NOISY_F32 = 7,
  MPC_SYN_MIXED_HASHED = 8, /* class = hash(seed, block) % 8 */
  MPC_SYN_MIXED_REGIONS = 9 /* class = block * 8 / total_blocks */
};
/* Fill d_lines (device, 128-byte blocks) with blocks [first_block, first_block + n_blocks) of the
 * synthetic dump `kind` with `seed`; total_blocks is the size of the whole (possibly sharded) dump. */
int mpc_synth_device(mpc_ctx* ctx, uint8_t* d_lines, uint64_t first_block, uint64_t n_blocks,
                     uint64_t total_blocks, int kind, uint64_t seed);

/* ---- stateless secondary compressors (BASELINE.json config #5) -------------------------------------------- */

enum { MPC_ALG_BDI = 1, MPC_ALG_FPC = 2, MPC_ALG_BPC = 3 };
/* counts: BDI = the nine BDIState counters (BDI.h:29-33); FPC = the eight per-word prefix counters (FPC.h:31-37;
 * TotalWords = their sum); BPC = the seven BPCPattern counters, counts[7] = TotalWords (BPC.h:29-33). */
typedef struct {
  uint64_t blocks, original_bits, compressed_bits;
  uint64_t counts[16];
} mpc_variant_stats;
/* One pass of comp::BDI / comp::FPC / comp::BPC::CompressLine (BDI.cpp:6, FPC.cpp:7, BPC.cpp:20) over n_blocks
 * 128-byte blocks.  sizes (optional): compressed bits per block.  FPC: the reference's zero-run scan reads past
 * the end of the line (FPC.cpp:26); here it stops at the block end -- sizes are identical, per-word statistics
 * are the deterministic ones. */
int mpc_variant_run_device(int alg, int device, const uint8_t* d_lines, uint64_t n_blocks, uint32_t line_size,
                           uint16_t* d_sizes, mpc_variant_stats* out, float* kernel_ms);
int mpc_variant_run_host(int alg, int device, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size,
                         uint16_t* h_sizes, mpc_variant_stats* out, float* kernel_ms);
const char* mpc_variant_error(void);

/* SC2 (Huffman over 32-bit words, SC2.cpp:270-334): device sort + run-length histogram of the first sampling_lines
 * lines, host tree with the reference's heap rules, device lookup.  sampling_lines is what main.cpp:108-114 derives
 * from the loader's row count: max(10000, min(rows / 100, 1000000)).  counts[0] = symbols that received a code. */
int mpc_sc2_run_device(int device, const uint8_t* d_lines, uint64_t n_blocks, uint32_t line_size, uint64_t sampling_lines,
                       uint16_t* d_sizes, mpc_variant_stats* out, float* kernel_ms);
int mpc_sc2_run_host(int device, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size, uint64_t sampling_lines,
                     uint16_t* h_sizes, mpc_variant_stats* out, float* kernel_ms);
const char* mpc_sc2_error(void);
/* The two phases on their own, for sharded runs: the GPU that holds the first sampling_lines lines
 * builds the code table, the table (at most 1024 x (symbol, length), plain host data) is handed to every shard -- by whatever
 * the job uses to broadcast -- and each shard applies it to its lines; first_block = index of the shard's first line in the
 * dump, so that the sampling lines (33 bits per word) are counted where they are. */
typedef struct {
  uint32_t n;              /* symbols that received a code (<= 1024) */
  uint32_t symbols[1024];  /* ascending */
  uint8_t lengths[1024];   /* Huffman code length = leaf depth, SC2.cpp:150-162 */
} mpc_sc2_table;
int mpc_sc2_build_table(int device, const uint8_t* d_lines, uint64_t sampling_lines, uint32_t line_size, mpc_sc2_table* table);
int mpc_sc2_apply_device(int device, const uint8_t* d_lines, uint64_t n_blocks, uint64_t first_block, uint64_t sampling_lines,
                         uint32_t line_size, const mpc_sc2_table* table, uint16_t* d_sizes, mpc_variant_stats* out, float* kernel_ms);

/* CPACK (CPACK.cpp:7-101): the 16-entry dictionary persists across lines, so the result depends on every earlier word
 * -- sequential by construction; this entry point runs on the host and is reported as such.
 * counts = ZZZZ, XXXX, MMMM, MMXX, ZZZX, MMMX (order of m_PatternLength, CPACK.h:119-127). */
int mpc_cpack_run_host(const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size, uint16_t* h_sizes,
                       mpc_variant_stats* out);

/* PATTERN (analysis tool, Pattern.cpp:6-75, Pattern.h:34-231): bytes of all-zero lines, of lines repeating one 4-byte
 * word, of lines seen before (temporal locality, LRU.h:17-56 with CACHESIZE = 2^24 - 1 lines), per base-delta layout
 * (B8D1, B8D2, B8D4, B4D1, B4D2, B2D1) the implicit (immediate) and explicit bytes of the lines that chose it, bytes of
 * lines no layout shrinks, and the byte histograms of all lines / of the lines that are neither all-zero nor
 * word-repeating (Pattern.h:96-125; the report derives two entropies from them).  Replaces comp::Pattern::CompressLine
 * + comp::PatternResult.  sizes (optional): the per-line return value (best layout + 4 bits).
 * cache_blocks: capacity of the temporal-locality cache in lines, 0 = the reference's 2^24 - 1.
 * temporal_path: 0 = device (hash sort; exact while the dump holds at most cache_blocks distinct lines), 1 = the cache
 * was simulated in order on the host (more distinct lines than the cache holds: evictions make it sequential). */
typedef struct {
  uint64_t blocks, total_bytes;
  uint64_t zeros_bytes, repeated_bytes, temporal_bytes, undefined_bytes;
  uint64_t implicit_bytes[6], explicit_bytes[6];
  uint64_t symbol_counts[256], symbol_counts_nontrivial[256];
  uint64_t distinct_blocks;
  int32_t temporal_path;
} mpc_pattern_stats;
int mpc_pattern_run_device(int device, const uint8_t* d_lines, uint64_t n_blocks, uint32_t line_size, uint64_t cache_blocks,
                           uint16_t* d_sizes, mpc_pattern_stats* out, float* kernel_ms);
int mpc_pattern_run_host(int device, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size, uint64_t cache_blocks,
                         uint16_t* h_sizes, mpc_pattern_stats* out, float* kernel_ms);
const char* mpc_pattern_error(void);

/* ---- library info ---------------------------------------------------------------------- */
const char* mpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MPC_CAPI_H_ */
