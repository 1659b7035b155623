// GPU size models of the stateless secondary compressors BDI / FPC / BPC (BASELINE.json config #5): one block per
// thread, same tile loader as the specialised MPC kernel.  Replaces comp::BDI/FPC/BPC::CompressLine
// (reference src/compressor/BDI.cpp:6, FPC.cpp:7, BPC.cpp:20) driven from main.cpp:237-243.
#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <string>

#include "mpc_capi.h"
#include "mpc_tile.cuh"
#include "mpc_variants.cuh"

namespace mpc {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kCounters = 16;  // device layout: [0] compressed bits, [1..16] counters
// Per algorithm: CTAs per SM the register allocation aims for and tile stages per warp (mpc_tile.cuh; one stage = the registers are the
// second buffer, 32 instead of 64 KiB of shared memory per CTA).  Measured on the 1 GiB hash-mixed / region-mixed dumps (GB/s):
//                 2 stages, 2 CTAs   2 stages, 3 CTAs   1 stage, 3 CTAs   1 stage, 4 CTAs
//   BDI (103 regs)   748 / 2 083        782 / 2 131        748 / 1 976       684 / 1 878
//   FPC (64 regs)  1 645 / 2 511      1 648 / 2 529      1 674 / 2 572     1 801 / 2 582
//   BPC (126 regs) 1 200 / 1 746      1 370 / 1 862      1 382 / 1 904     1 281 / 1 644
template <int ALG> struct VariantTune { static constexpr int kMinCtas = 3, kStages = 2; };           // BDI
template <> struct VariantTune<MPC_ALG_FPC> { static constexpr int kMinCtas = 4, kStages = 1; };
template <> struct VariantTune<MPC_ALG_BPC> { static constexpr int kMinCtas = 3, kStages = 1; };

// warp-wide votes for mpcvar::bdi_block_with: every lane calls the maker, the lanes that run the checks vote among themselves
struct WarpVote {
  unsigned mask;
  __device__ __forceinline__ bool operator()(bool b) const { return __all_sync(mask, b) != 0; }
};
struct MakeWarpVote {
  __device__ __forceinline__ WarpVote operator()(bool runs_checks) const { return WarpVote{__ballot_sync(0xffffffffu, runs_checks)}; }
};

// W = words per line: a thread holds 128 bytes = 32 / W consecutive lines (32-, 64- and 128-byte lines; the reference's
// models take any line size, BDI.cpp:108-201, FPC.cpp:7-87, BPC.cpp:20-185)
template <int ALG, int W>
__global__ void __launch_bounds__(kThreads, VariantTune<ALG>::kMinCtas)
variant_kernel(const uint4* __restrict__ lines, uint64_t n_blocks, uint16_t* __restrict__ sizes,
               unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* s_stage = reinterpret_cast<uint4*>(smem_raw);
  __shared__ unsigned long long s_cnt[1 + kCounters];
  // per-tile counters: one row per warp, written by that warp's lane 0 with plain loads and stores.  64-bit shared-memory atomics
  // are compare-and-swap loops (ATOMS.CAST.SPIN.64), and with every warp of the CTA on the same few words a finely mixed dump -- up
  // to nine counters per tile -- spent more time spinning than checking.  A row counts the lines / words of its warp's tiles: far
  // below 2^32 for any dump one launch admits.
  __shared__ uint32_t s_wcnt[kWarps][kCounters];
  if (threadIdx.x < 1 + kCounters) s_cnt[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < kWarps * kCounters; i += kThreads) (&s_wcnt[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long bits = 0;
  // FPC (64 registers: room to spare): per-thread counters -- a block's eight 8-bit counters (<= 32 each) are added into one 64-bit
  // word, which is unpacked into 32-bit registers every seventh block (7 x 32 < 256): a couple of instructions per block instead of eight
  // warp reductions and atomics, the warp reductions happen once after the last tile (+7 % on the hash-mixed dump).  BDI and BPC sit at
  // 103 / 126 registers: the ten extra live registers spill there (measured -17 % / -29 %), so they aggregate per tile as before.
  constexpr bool kThreadCounters = ALG == MPC_ALG_FPC;
  unsigned long long acc8 = 0;
  uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, in_acc8 = 0;
  auto unpack = [&]() {
#pragma unroll
    for (int p = 0; p < 8; p++) acc[p] += (uint32_t)(acc8 >> (8 * p)) & 0xffu;
    acc8 = 0;
    in_acc8 = 0;
  };
  constexpr int S = 32 / W;  // lines per thread
  // the tile loader works on 128-byte units: n_blocks lines = ceil(n_blocks / S) units, the last one possibly partial
  constexpr int kVariantStages = VariantTune<ALG>::kStages;
  tile::for_each_block<kVariantStages>(lines, (n_blocks + S - 1) / S, s_stage + warp * kVariantStages * 256, kWarps,
                       [&](const uint32_t (&x128)[32], uint64_t unit, bool) {
#pragma unroll
   for (int sub = 0; sub < S; sub++) {
    uint32_t x[32];
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = x128[sub * W + i];
    const uint64_t blk = unit * S + sub;
    const bool valid = blk < n_blocks;
    uint32_t size = 0;
    uint64_t packed = 0;  // eight 8-bit counters
    uint32_t extra = 0;
    if (ALG == MPC_ALG_BDI) {
      int st;
      size = mpcvar::bdi_block_with<W>(x, &st, MakeWarpVote());
      packed = 1ull << (4 * st);  // nine 4-bit one-hot counters
    } else if (ALG == MPC_ALG_FPC) {
      size = mpcvar::fpc_block<W>(x, &packed, WarpVote{0xffffffffu});  // every lane is here
    } else {
      size = mpcvar::bpc_block<W>(x, &packed, &extra);
    }
    if (!valid) { size = 0; packed = 0; extra = 0; }
    if (valid && sizes) sizes[blk] = (uint16_t)size;
    bits += size;
    if (kThreadCounters) {
      acc8 += packed;
      if (++in_acc8 == 7) unpack();
    } else if (ALG == MPC_ALG_BDI) {  // warp-aggregate the small counters, one shared atomic per counter per tile
#pragma unroll
      for (int s = 0; s < 9; s++) {
        const uint32_t c = __popc(__ballot_sync(0xffffffffu, (packed >> (4 * s)) & 1ull));
        if (lane == 0 && c) s_wcnt[warp][s] += c;
      }
    } else {
#pragma unroll
      for (int p = 0; p < 8; p++) {
        const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)((packed >> (8 * p)) & 0xffull));
        if (lane == 0 && c) s_wcnt[warp][p] += c;
      }
      const uint32_t c = __reduce_add_sync(0xffffffffu, extra);
      if (lane == 0 && c) s_wcnt[warp][7] += c;  // counts[7] = TotalWords
    }
   }  // sub-lines
  }, n_blocks * (uint64_t)(W / 4));
  if (kThreadCounters) {
    unpack();
#pragma unroll
    for (int p = 0; p < 8; p++) {
      const uint32_t c = __reduce_add_sync(0xffffffffu, acc[p]);
      if (lane == 0 && c) atomicAdd(&s_cnt[1 + p], (unsigned long long)c);
    }
  }
  // per-thread bit totals -> warp -> CTA -> global
  for (int o = 16; o; o >>= 1) bits += __shfl_down_sync(0xffffffffu, bits, o);
  if (lane == 0 && bits) atomicAdd(&s_cnt[0], bits);
  __syncthreads();
  if (threadIdx.x < 1 + kCounters) {
    unsigned long long v = s_cnt[threadIdx.x];
    if (threadIdx.x >= 1)
      for (int w = 0; w < kWarps; w++) v += s_wcnt[w][threadIdx.x - 1];
    if (v) atomicAdd(&stats[threadIdx.x], v);
  }
}

thread_local std::string g_verr;

int vfail(int code, const char* what, cudaError_t e) {
  g_verr = std::string(what) + ": " + cudaGetErrorString(e);
  return code;
}

template <int ALG, int W>
cudaError_t launch_variant_w(const uint8_t* d_lines, uint64_t n, uint16_t* d_sizes, unsigned long long* d_stats, int sms,
                             cudaStream_t s) {
  const size_t smem = (size_t)kWarps * VariantTune<ALG>::kStages * tile::kTileBytes;
  cudaError_t e = cudaFuncSetAttribute(variant_kernel<ALG, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, variant_kernel<ALG, W>, kThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const uint64_t units = (n + (32 / W) - 1) / (32 / W);  // 128-byte units
  const uint64_t tiles = (units + tile::kTileBlocks - 1) / tile::kTileBlocks;
  uint64_t grid = (uint64_t)sms * per_sm;
  const uint64_t want = (tiles + kWarps - 1) / kWarps;
  if (grid > want) grid = want;
  variant_kernel<ALG, W><<<(unsigned)grid, kThreads, smem, s>>>(reinterpret_cast<const uint4*>(d_lines), n, d_sizes, d_stats);
  return cudaGetLastError();
}
template <int ALG>
cudaError_t launch_variant(const uint8_t* d_lines, uint64_t n, uint32_t line_size, uint16_t* d_sizes, unsigned long long* d_stats, int sms,
                           cudaStream_t s) {
  if (line_size == 32) return launch_variant_w<ALG, 8>(d_lines, n, d_sizes, d_stats, sms, s);
  if (line_size == 64) return launch_variant_w<ALG, 16>(d_lines, n, d_sizes, d_stats, sms, s);
  return launch_variant_w<ALG, 32>(d_lines, n, d_sizes, d_stats, sms, s);
}

}  // namespace
}  // namespace mpc

extern "C" const char* mpc_variant_error(void) { return mpc::g_verr.c_str(); }

namespace mpc {
namespace {
// per-device counters + timing events, created once per process (no allocation or event creation per call)
struct VariantWorkspace {
  bool ready = false;
  unsigned long long* d_stats = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};
std::mutex g_vws_mutex;
VariantWorkspace g_vws[16];
}  // namespace
}  // namespace mpc

extern "C" int mpc_variant_run_device(int alg, int device, const uint8_t* d_lines, uint64_t n_blocks, uint32_t line_size,
                                      uint16_t* d_sizes, mpc_variant_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (!out || (n_blocks && !d_lines)) { g_verr = "null argument"; return MPC_E_ARG; }
  if (alg < MPC_ALG_BDI || alg > MPC_ALG_BPC) { g_verr = "unknown algorithm id"; return MPC_E_ARG; }
  if (line_size != 32 && line_size != 64 && line_size != 128) { g_verr = "line size must be 32, 64 or 128 bytes"; return MPC_E_ARG; }
  if ((uintptr_t)d_lines & 15) { g_verr = "lines must be 16-byte aligned"; return MPC_E_ARG; }
  if (device < 0 || device >= 16) { g_verr = "device index out of range"; return MPC_E_ARG; }
  std::lock_guard<std::mutex> lock(g_vws_mutex);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return vfail(MPC_E_CUDA, "cudaSetDevice", e);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  VariantWorkspace& w = g_vws[device];
  if (!w.ready) {
    if ((e = cudaMalloc(&w.d_stats, (1 + kCounters) * sizeof(unsigned long long))) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaMalloc", e);
    if ((e = cudaEventCreate(&w.e0)) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaEventCreate", e);
    if ((e = cudaEventCreate(&w.e1)) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaEventCreate", e);
    w.ready = true;
  }
  if ((e = cudaMemsetAsync(w.d_stats, 0, (1 + kCounters) * sizeof(unsigned long long), 0)) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaMemsetAsync", e);
  if ((e = cudaEventRecord(w.e0, 0)) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaEventRecord", e);
  if (n_blocks) {
    if (alg == MPC_ALG_BDI) e = launch_variant<MPC_ALG_BDI>(d_lines, n_blocks, line_size, d_sizes, w.d_stats, sms, 0);
    else if (alg == MPC_ALG_FPC) e = launch_variant<MPC_ALG_FPC>(d_lines, n_blocks, line_size, d_sizes, w.d_stats, sms, 0);
    else e = launch_variant<MPC_ALG_BPC>(d_lines, n_blocks, line_size, d_sizes, w.d_stats, sms, 0);
    if (e != cudaSuccess) return vfail(MPC_E_CUDA, "variant kernel launch", e);
  }
  if ((e = cudaEventRecord(w.e1, 0)) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaEventRecord", e);
  unsigned long long h[1 + kCounters];
  if ((e = cudaMemcpy(h, w.d_stats, sizeof(h), cudaMemcpyDeviceToHost)) != cudaSuccess) return vfail(MPC_E_CUDA, "variant kernel", e);
  float ms = 0.f;
  if ((e = cudaEventElapsedTime(&ms, w.e0, w.e1)) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaEventElapsedTime", e);
  memset(out, 0, sizeof(*out));
  out->blocks = n_blocks;
  out->original_bits = n_blocks * 8ull * line_size;  // CompResult::OriginalSize
  out->compressed_bits = h[0];                       // CompResult::CompressedSize
  for (int i = 0; i < kCounters; i++) out->counts[i] = h[1 + i];
  if (kernel_ms) *kernel_ms = ms;
  return MPC_OK;
}

// Host dump: streamed through one device buffer in 256 MiB chunks (dumps larger than HBM work), every copy checked.
extern "C" int mpc_variant_run_host(int alg, int device, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size,
                                    uint16_t* h_sizes, mpc_variant_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (!out || (n_blocks && !h_lines)) { g_verr = "null argument"; return MPC_E_ARG; }
  if (line_size != 32 && line_size != 64 && line_size != 128) { g_verr = "line size must be 32, 64 or 128 bytes"; return MPC_E_ARG; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return vfail(MPC_E_CUDA, "cudaSetDevice", e);
  const uint64_t chunk = (256ull << 20) / line_size;
  const uint64_t cap = n_blocks < chunk ? (n_blocks ? n_blocks : 1) : chunk;
  uint8_t* d_lines = nullptr;
  uint16_t* d_sizes = nullptr;
  if ((e = cudaMalloc(&d_lines, (size_t)cap * line_size)) != cudaSuccess) return vfail(MPC_E_CUDA, "cudaMalloc", e);
  if (h_sizes && (e = cudaMalloc(&d_sizes, (size_t)cap * sizeof(uint16_t))) != cudaSuccess) {
    cudaFree(d_lines);
    return vfail(MPC_E_CUDA, "cudaMalloc", e);
  }
  auto done = [&](int rc) { cudaFree(d_lines); if (d_sizes) cudaFree(d_sizes); return rc; };
  memset(out, 0, sizeof(*out));
  float ms_total = 0.f;
  for (uint64_t lo = 0; lo < n_blocks; lo += chunk) {
    const uint64_t nb = n_blocks - lo < chunk ? n_blocks - lo : chunk;
    if ((e = cudaMemcpy(d_lines, h_lines + lo * line_size, (size_t)nb * line_size, cudaMemcpyHostToDevice)) != cudaSuccess) return done(vfail(MPC_E_CUDA, "H2D copy", e));
    mpc_variant_stats part;
    float ms = 0.f;
    const int rc = mpc_variant_run_device(alg, device, d_lines, nb, line_size, d_sizes, &part, &ms);
    if (rc != MPC_OK) return done(rc);
    if (h_sizes && (e = cudaMemcpy(h_sizes + lo, d_sizes, (size_t)nb * sizeof(uint16_t), cudaMemcpyDeviceToHost)) != cudaSuccess) return done(vfail(MPC_E_CUDA, "D2H copy", e));
    out->blocks += part.blocks;
    out->original_bits += part.original_bits;
    out->compressed_bits += part.compressed_bits;
    for (int i = 0; i < 16; i++) out->counts[i] += part.counts[i];
    ms_total += ms;
  }
  if (kernel_ms) *kernel_ms = ms_total;
  return done(MPC_OK);
}
