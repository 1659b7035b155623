// PATTERN -- the reference's data-pattern analysis (src/compressor/Pattern.{h,cpp}, LRU.h) behind the C ABI.
//
// Per line (Pattern.cpp:6-75): zero / 4-byte-repeat flags, the best of six base-delta layouts (checkPattern ==
// BDI's checkBDI) with its implicit (immediate) and explicit byte counts (countPattern, Pattern.cpp:201-320), and two
// byte histograms (all lines; lines that are neither all-zero nor word-repeating, Pattern.h:96-125) from which the
// report computes two entropies.  One block per thread through the tile loader shared with the other variants.
//
// Byte histogram of the non-trivial lines: 128 increments per line.  Shared-memory atomics cost two cycles per LANE on this
// hardware whatever the addresses are (64 cycles per warp instruction: 8 192 cycles per tile of 32 lines, the whole kernel at
// 0.1-0.2 TB/s), plain shared-memory loads and stores one cycle per conflict-free WARP instruction.  So every lane keeps its own
// counters and updates them without atomics: a warp owns an 8 KiB table of 8-bit counters, counter (byte value v, lane l) at
// byte 32 v + l -- no two lanes ever touch the same counter, and only lanes of the same quad can meet in a bank.  A counter that
// is about to wrap passes 256 on to the CTA's 32-bit table (one atomic per 256 increments); the warp sums its table into
// that table when it has run out of tiles.
//
// Temporal locality (Pattern.cpp:101-107, LRU.h:17-56): a line counts when an identical line is in the cache.  The
// reference never promotes a hit and inserts only on a miss, so the cache is a FIFO set of the last C = 2^24 - 1
// distinct lines that missed.  While the dump holds at most C distinct lines nothing is ever evicted and the count is
// simply (lines - distinct lines): the kernel writes a 64-bit content hash per line, CUB sorts (hash, index) and a
// second kernel confirms every equal-hash neighbour pair on the 128 bytes themselves.  A hash collision between
// different lines, or more than C distinct lines, sends the pass to the host, where the FIFO is simulated in order
// (sequential by construction, like CPACK's dictionary).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "mpc_capi.h"
#include "mpc_tile.cuh"
#include "mpc_variants.cuh"

namespace mpc {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kPatStages = 1;             // tile stages per warp (the registers are the second buffer)
constexpr uint32_t kLaneTableBytes = 256 * 32;  // per warp: one 8-bit counter per (byte value, lane)
// device counters: [0] zero lines [1] repeat lines [2] undefined lines [3..8] lines per pattern [9..14] immediates per
// pattern [15] duplicate lines [16] hash collisions ; then two 256-bin byte histograms (trivial lines / other lines)
constexpr int kCnt = 17;
constexpr int kWords = kCnt + 512;

thread_local std::string g_perr;
int pfail(int code, const std::string& what) { g_perr = what; return code; }
#define PAT_CUDA(call)                                                                                     \
  do {                                                                                                     \
    cudaError_t e__ = (call);                                                                              \
    if (e__ != cudaSuccess) return pfail(MPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// The four bytes of one line word into this lane's 8-bit counters.  `table` = shared-memory address of the warp's table (8 KiB
// aligned) + lane, so that a counter's address is one LOP3: table | 32 * byte value, the latter being the word shifted so that the
// byte sits at bits 5..12 (shifts on the FMA pipe).  The loads and stores of equal bytes hit the same counter and stay in program
// order; a counter that wrapped to 0 (one test per word) hands 256 to the CTA's 32-bit bin of its byte value.
__device__ __forceinline__ uint32_t lane_count(uint32_t addr) {
  uint32_t c;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(addr) : "memory");
  c += 1u;
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(c) : "memory");
  return c;
}
#ifndef MPC_PAT_LAYOUT
#define MPC_PAT_LAYOUT 0
#endif
// shared-memory address of this lane's counter of byte k of the word v; bin = the byte value the address stands for
#if MPC_PAT_LAYOUT == 0
// counter (value b, lane l) at byte 32 b + l: one shift (FMA pipe) and one LOP3 per byte; lanes of one quad that hold different
// values with equal b mod 4 meet in a bank (2- to 4-way on random bytes, none on the equal bytes of structured data)
template <int K>
__device__ __forceinline__ uint32_t lane_counter(uint32_t v, uint32_t table) {
  const uint32_t s = K == 0 ? (v << 5) : mpcdev::shr_fma(v, 8 * K - 5);
  return table | (s & 0x1fe0u);
}
__device__ __forceinline__ uint32_t counter_bin(uint32_t addr) { return (addr >> 5) & 0xffu; }
#else
// counter (value b, lane l) at byte 128 (b >> 2) + 4 l + (b & 3): the bank is the lane whatever the values are, at the price of a
// second shift and LOP3 per byte.  `table` = table base + 4 * lane.
template <int K>
__device__ __forceinline__ uint32_t lane_counter(uint32_t v, uint32_t table) {
  const uint32_t s = K == 0 ? (v << 5) : mpcdev::shr_fma(v, 8 * K - 5);
  const uint32_t t = K == 0 ? v : mpcdev::shr_fma(v, 8 * K);
  return (s & 0x1f80u) | ((t & 3u) | table);
}
__device__ __forceinline__ uint32_t counter_bin(uint32_t addr) { return ((addr >> 5) & 0xfcu) | (addr & 3u); }
#endif
__device__ __forceinline__ void lane_count_word(uint32_t v, uint32_t table, uint32_t* s_bins) {
  const uint32_t a0 = lane_counter<0>(v, table), a1 = lane_counter<1>(v, table), a2 = lane_counter<2>(v, table), a3 = lane_counter<3>(v, table);
  const uint32_t c0 = lane_count(a0), c1 = lane_count(a1), c2 = lane_count(a2), c3 = lane_count(a3);
  if ((c0 | c1 | c2 | c3) & 0x100u) {
    if (c0 == 256u) atomicAdd(&s_bins[counter_bin(a0)], 256u);
    if (c1 == 256u) atomicAdd(&s_bins[counter_bin(a1)], 256u);
    if (c2 == 256u) atomicAdd(&s_bins[counter_bin(a2)], 256u);
    if (c3 == 256u) atomicAdd(&s_bins[counter_bin(a3)], 256u);
  }
}

// W = words per line: a thread holds 128 bytes = 32 / W consecutive lines (the analysis takes any line size, Pattern.cpp:6-75)
template <int W>
__global__ void __launch_bounds__(kThreads)
pattern_kernel(const uint4* __restrict__ lines, uint64_t n_blocks, uint16_t* __restrict__ sizes, uint64_t* __restrict__ hashes,
               unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* s_stage = reinterpret_cast<uint4*>(smem_raw);
  // 32-bit words: a 64-bit shared-memory atomicAdd is a compare-and-swap loop, and every warp of the CTA adds to these few words once per
  // tile.  A CTA sees at most ceil(lines / CTAs) lines and 64 immediates per line: below 2^32 for any dump the 32-bit line index admits
  // (a grid smaller than the GPU means a dump of a few tiles).
  __shared__ uint32_t s_cnt[kCnt];
  __shared__ uint32_t s_hist[512];  // [0..255] bytes of all-zero / word-repeating lines, [256..511] bytes of the others
  if (threadIdx.x < kCnt) s_cnt[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < 512; i += kThreads) s_hist[i] = 0;
  // per-warp tables of per-lane 8-bit byte counters behind the tile stages, 8 KiB aligned in the shared-memory window
  const uint32_t stages_end = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)(kWarps * kPatStages * tile::kTileBytes);
  const uint32_t tables_at = (stages_end + (kLaneTableBytes - 1u)) & ~(kLaneTableBytes - 1u);
  uint32_t* s_tables = reinterpret_cast<uint32_t*>(smem_raw + (tables_at - (uint32_t)__cvta_generic_to_shared(smem_raw)));
  for (int i = threadIdx.x; i < kWarps * (kLaneTableBytes / 4); i += kThreads) s_tables[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t my_table = tables_at + (uint32_t)warp * kLaneTableBytes + (uint32_t)lane * (MPC_PAT_LAYOUT == 0 ? 1u : 4u);
  uint32_t* s_bins = s_hist + 256;
  constexpr int S = 32 / W;
  tile::for_each_block<kPatStages>(lines, (n_blocks + S - 1) / S, s_stage + warp * kPatStages * 256, kWarps,
                       [&](const uint32_t (&x128)[32], uint64_t unit, bool) {
#pragma unroll
   for (int sub = 0; sub < S; sub++) {
    uint32_t x[32];
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = x128[sub * W + i];
    const uint64_t blk = unit * S + sub;
    const bool valid = blk < n_blocks;
    uint32_t any = 0, rep = 0;
#pragma unroll
    for (int i = 0; i < W; i++) { any |= x[i]; rep |= x[i] ^ x[0]; }
    const bool is_zero = valid && any == 0;   // Pattern::isZeros
    const bool is_rep = valid && rep == 0;    // Pattern::isRepeated(line, 4) == PatternResult::IsAllWordSame
    int sel;
    uint32_t imm;
    struct { __device__ __forceinline__ bool operator()(bool b) const { return __all_sync(0xffffffffu, b) != 0; } } vote;  // every lane is here
    const uint32_t size = mpcvar::pattern_block<W>(x, &sel, &imm, vote);
    if (valid && sizes) sizes[blk] = (uint16_t)size;
    if (valid && hashes) hashes[blk] = mpcvar::block_hash64<W>(x);
    // byte histograms (PatternResult::UpdateCountMap): a word-repeating line is W copies of its first word
    if (is_rep) {
#pragma unroll
      for (int k = 0; k < 4; k++) atomicAdd(&s_hist[(x[0] >> (8 * k)) & 0xffu], (uint32_t)W);
    } else if (valid) {
      // a copy of the line that the loop indexes at run time (thread-local memory, L1-resident): the 128 updates are eight iterations
      // of a four-word body instead of 700 straight-line instructions -- the kernel's footprint was past the instruction cache
      // (profiles/r02_end_pattern_smooth.txt: more than half of the stall samples were instruction fetches).  Measured per GiB:
      // unrolled 2.91 ms (smooth) / 3.23 (hash-mixed), one word per iteration 2.35 / 3.70, four 2.00 / 3.34, eight 2.11 / 3.42.
      uint32_t xl[W];
#pragma unroll
      for (int i = 0; i < W; i++) xl[i] = x[i];
#pragma unroll 4
      for (int i = 0; i < W; i++) lane_count_word(xl[i], my_table, s_bins);
    }
    // warp-aggregated line counters
    const uint32_t cz = __popc(__ballot_sync(0xffffffffu, is_zero)), cr = __popc(__ballot_sync(0xffffffffu, is_rep));
    const uint32_t cu = __popc(__ballot_sync(0xffffffffu, valid && sel == 9));
    if (lane == 0) {
      if (cz) atomicAdd(&s_cnt[0], cz);
      if (cr) atomicAdd(&s_cnt[1], cr);
      if (cu) atomicAdd(&s_cnt[2], cu);
    }
#pragma unroll
    for (int p = 0; p < 6; p++) {
      const bool mine = valid && sel == p;
      const uint32_t c = __popc(__ballot_sync(0xffffffffu, mine));
      if (c) {  // warp-uniform
        const uint32_t s = __reduce_add_sync(0xffffffffu, mine ? imm : 0u);
        if (lane == 0) { atomicAdd(&s_cnt[3 + p], c); atomicAdd(&s_cnt[9 + p], s); }
      }
    }
   }  // sub-lines
  }, n_blocks * (uint64_t)(W / 4));
  // this warp's counters into the CTA's bins: lane l sums the 32 lane counters (8 words) of values l, l + 32, ...
  __syncwarp();
  {
    const uint32_t* t = s_tables + warp * (kLaneTableBytes / 4);
#pragma unroll 1
    for (int v = lane; v < 256; v += 32) {
      uint32_t sum = 0;
#if MPC_PAT_LAYOUT == 0
#pragma unroll
      for (int j = 0; j < 8; j++) sum = __dp4a(t[v * 8 + ((j + (lane >> 2)) & 7)], 0x01010101u, sum);
#else
#pragma unroll 8
      for (int j = 0; j < 32; j++) sum += (t[(v >> 2) * 32 + ((j + (lane >> 2)) & 31)] >> (8 * (v & 3))) & 0xffu;
#endif
      if (sum) atomicAdd(&s_bins[v], sum);
    }
  }
  __syncthreads();
  if (threadIdx.x < kCnt && s_cnt[threadIdx.x]) atomicAdd(&stats[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
  for (int i = threadIdx.x; i < 512; i += kThreads)
    if (s_hist[i]) atomicAdd(&stats[kCnt + i], (unsigned long long)s_hist[i]);
}

__global__ void iota_kernel(uint32_t* __restrict__ v, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// sorted (hash, index) pairs: a line whose hash equals its predecessor's is a duplicate when the 128 bytes agree, and
// a hash collision otherwise (the caller then falls back to the exact host pass)
__global__ void dup_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx, const uint4* __restrict__ lines,
                           uint32_t n, uint32_t chunks, unsigned long long* __restrict__ stats) {  // chunks = line size / 16
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  bool dup = false, coll = false;
  if (j >= 1 && j < n && keys[j] == keys[j - 1]) {
    const uint4* a = lines + (uint64_t)idx[j] * chunks;
    const uint4* b = lines + (uint64_t)idx[j - 1] * chunks;
    uint32_t d = 0;
    for (uint32_t k = 0; k < chunks; k++) {
      const uint4 p = a[k], q = b[k];
      d |= (p.x ^ q.x) | (p.y ^ q.y) | (p.z ^ q.z) | (p.w ^ q.w);
    }
    dup = d == 0;
    coll = d != 0;
  }
  const uint32_t cd = __popc(__ballot_sync(0xffffffffu, dup)), cc = __popc(__ballot_sync(0xffffffffu, coll));
  if ((threadIdx.x & 31) == 0) {
    if (cd) atomicAdd(&stats[15], (unsigned long long)cd);
    if (cc) atomicAdd(&stats[16], (unsigned long long)cc);
  }
}

// Exact, sequential simulation of the reference cache (LRU.h:17-56 as Pattern.cpp:101-107 drives it): returns the
// number of lines that hit.  Open-addressing table of line indices; evicted entries become tombstones.
uint64_t temporal_hits_host(const uint8_t* lines, uint64_t n, uint32_t L, uint64_t capacity) {
  auto hash = [&](const uint8_t* p) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < L; i += 8) {
      uint64_t v;
      memcpy(&v, p + i, 8);
      h = (h ^ v) * 0xFF51AFD7ED558CCDull;
      h ^= h >> 29;
    }
    return h;
  };
  uint64_t slots = 64;
  const uint64_t live_max = n < capacity ? n : capacity;
  while (slots < 4 * live_max + 64) slots <<= 1;
  const uint64_t kEmpty = 0, kTomb = ~0ull;
  std::vector<uint64_t> table(slots, kEmpty);
  std::vector<uint64_t> order;  // insertion order of the misses
  order.reserve(n);
  uint64_t head = 0, live = 0, used = 0, hits = 0;
  for (uint64_t i = 0; i < n; i++) {
    const uint8_t* line = lines + i * L;
    uint64_t pos = hash(line) & (slots - 1), tomb = kTomb;
    bool hit = false;
    while (table[pos] != kEmpty) {
      if (table[pos] == kTomb) { if (tomb == kTomb) tomb = pos; }
      else if (memcmp(lines + (table[pos] - 1) * L, line, L) == 0) { hit = true; break; }
      pos = (pos + 1) & (slots - 1);
    }
    if (hit) { hits++; continue; }
    if (tomb != kTomb) pos = tomb; else used++;
    table[pos] = i + 1;
    order.push_back(i);
    live++;
    while (live > capacity) {  // LRUCache::clean
      const uint64_t victim = order[head++];
      uint64_t p = hash(lines + victim * L) & (slots - 1);
      while (table[p] != victim + 1) p = (p + 1) & (slots - 1);
      table[p] = kTomb;
      live--;
    }
    if (used * 2 > slots) {  // too many tombstones: rebuild from the live entries
      std::fill(table.begin(), table.end(), kEmpty);
      used = 0;
      for (uint64_t q = head; q < order.size(); q++) {
        uint64_t p = hash(lines + order[q] * L) & (slots - 1);
        while (table[p] != kEmpty) p = (p + 1) & (slots - 1);
        table[p] = order[q] + 1;
        used++;
      }
    }
  }
  return hits;
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

}  // namespace
}  // namespace mpc

extern "C" const char* mpc_pattern_error(void) { return mpc::g_perr.c_str(); }

// h_lines: host copy of the same lines when the caller has one (used only if the temporal pass must run on the host)
static int pattern_run(int device, const uint8_t* d_lines, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size,
                       uint64_t cache_blocks, uint16_t* d_sizes, mpc_pattern_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (!out || (n_blocks && !d_lines)) return pfail(MPC_E_ARG, "null argument");
  if (line_size != 32 && line_size != 64 && line_size != 128) return pfail(MPC_E_ARG, "line size must be 32, 64 or 128 bytes");
  if ((uintptr_t)d_lines & 15) return pfail(MPC_E_ARG, "lines must be 16-byte aligned");
  if (n_blocks >= (1ull << 31)) return pfail(MPC_E_ARG, "too many blocks for 32-bit line indices");
  if (cache_blocks == 0) cache_blocks = (1ull << 24) - 1;  // CACHESIZE, LRU.h:6
  PAT_CUDA(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const uint32_t n = (uint32_t)n_blocks;
  DevBuf stats, keys, keys2, idx, idx2, temp;
  PAT_CUDA(cudaMalloc(&stats.p, kWords * sizeof(unsigned long long)));
  PAT_CUDA(cudaMemset(stats.p, 0, kWords * sizeof(unsigned long long)));
  PAT_CUDA(cudaMalloc(&keys.p, (size_t)(n ? n : 1) * 8));
  // every buffer of the sort is allocated BEFORE the first kernel: cudaMalloc between two launches leaves the GPU idle for as long as
  // the allocation takes on the host (milliseconds in a process that already holds most of the memory), inside the timed region
  size_t tbytes = 0;
  if (n) {
    PAT_CUDA(cudaMalloc(&keys2.p, (size_t)n * 8));
    PAT_CUDA(cudaMalloc(&idx.p, (size_t)n * 4));
    PAT_CUDA(cudaMalloc(&idx2.p, (size_t)n * 4));
    PAT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tbytes, (const uint64_t*)keys.p, (uint64_t*)keys2.p, (const uint32_t*)idx.p,
                                             (uint32_t*)idx2.p, (int)n));
    PAT_CUDA(cudaMalloc(&temp.p, tbytes ? tbytes : 16));
  }
  EventPair ev;
  cudaEventRecord(ev.a, 0);
  if (n) {
    const size_t smem = (size_t)kWarps * kPatStages * tile::kTileBytes + (size_t)(kWarps + 1) * kLaneTableBytes;  // + alignment slack
    auto launch = [&](auto kernel, uint64_t lines_per_unit) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      int per_sm = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
      if (e != cudaSuccess) return e;
      if (per_sm < 1) per_sm = 1;
      const uint64_t units = (n_blocks + lines_per_unit - 1) / lines_per_unit;  // 128-byte units
      const uint64_t tiles = (units + tile::kTileBlocks - 1) / tile::kTileBlocks;
      uint64_t grid = (uint64_t)sms * per_sm;
      const uint64_t want = (tiles + kWarps - 1) / kWarps;
      if (grid > want) grid = want;
      kernel<<<(unsigned)grid, kThreads, smem>>>(reinterpret_cast<const uint4*>(d_lines), n_blocks, d_sizes, (uint64_t*)keys.p,
                                                 (unsigned long long*)stats.p);
      return cudaGetLastError();
    };
    if (line_size == 32) PAT_CUDA(launch(pattern_kernel<8>, 4));
    else if (line_size == 64) PAT_CUDA(launch(pattern_kernel<16>, 2));
    else PAT_CUDA(launch(pattern_kernel<32>, 1));
    // temporal locality: sort (hash, index), confirm equal-hash neighbours on the bytes
    iota_kernel<<<(n + 255) / 256, 256>>>((uint32_t*)idx.p, n);
    PAT_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tbytes, (const uint64_t*)keys.p, (uint64_t*)keys2.p, (const uint32_t*)idx.p,
                                             (uint32_t*)idx2.p, (int)n));
    dup_kernel<<<(n + 255) / 256, 256>>>((const uint64_t*)keys2.p, (const uint32_t*)idx2.p, reinterpret_cast<const uint4*>(d_lines), n,
                                         line_size / 16, (unsigned long long*)stats.p);
    PAT_CUDA(cudaGetLastError());
  }
  cudaEventRecord(ev.b, 0);
  std::vector<unsigned long long> h(kWords);
  PAT_CUDA(cudaMemcpy(h.data(), stats.p, kWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev.a, ev.b);

  memset(out, 0, sizeof(*out));
  static const uint32_t kBase[6] = {8, 8, 8, 4, 4, 2};
  out->blocks = n_blocks;
  out->total_bytes = n_blocks * line_size;
  out->zeros_bytes = h[0] * line_size;
  out->repeated_bytes = h[1] * line_size;
  out->undefined_bytes = h[2] * line_size;
  for (int p = 0; p < 6; p++) {
    const uint64_t values = h[3 + p] * (line_size / kBase[p]);
    out->implicit_bytes[p] = h[9 + p] * kBase[p];
    out->explicit_bytes[p] = (values - h[9 + p]) * kBase[p];
  }
  for (int b = 0; b < 256; b++) {
    out->symbol_counts[b] = h[kCnt + b] + h[kCnt + 256 + b];
    out->symbol_counts_nontrivial[b] = h[kCnt + 256 + b];
  }
  const uint64_t dups = h[15], collisions = h[16];
  out->distinct_blocks = n_blocks - dups;
  if (collisions == 0 && n_blocks - dups <= cache_blocks) {
    out->temporal_bytes = dups * line_size;
    out->temporal_path = 0;
  } else {
    std::vector<uint8_t> copy;
    if (!h_lines) {
      copy.resize((size_t)n_blocks * line_size);
      PAT_CUDA(cudaMemcpy(copy.data(), d_lines, copy.size(), cudaMemcpyDeviceToHost));
      h_lines = copy.data();
    }
    out->temporal_bytes = temporal_hits_host(h_lines, n_blocks, line_size, cache_blocks) * line_size;
    out->temporal_path = 1;
    if (collisions) out->distinct_blocks = 0;  // unknown: equal hashes with different bytes
  }
  if (kernel_ms) *kernel_ms = ms;
  return MPC_OK;
}

extern "C" int mpc_pattern_run_device(int device, const uint8_t* d_lines, uint64_t n_blocks, uint32_t line_size,
                                      uint64_t cache_blocks, uint16_t* d_sizes, mpc_pattern_stats* out, float* kernel_ms) {
  return pattern_run(device, d_lines, nullptr, n_blocks, line_size, cache_blocks, d_sizes, out, kernel_ms);
}

extern "C" int mpc_pattern_run_host(int device, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size,
                                    uint64_t cache_blocks, uint16_t* h_sizes, mpc_pattern_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (n_blocks && !h_lines) return pfail(MPC_E_ARG, "null lines");
  PAT_CUDA(cudaSetDevice(device));
  DevBuf lines, sizes;
  const size_t bytes = (size_t)n_blocks * line_size;
  PAT_CUDA(cudaMalloc(&lines.p, bytes ? bytes : 16));
  if (h_sizes) PAT_CUDA(cudaMalloc(&sizes.p, (n_blocks ? n_blocks : 1) * sizeof(uint16_t)));
  if (bytes) PAT_CUDA(cudaMemcpy(lines.p, h_lines, bytes, cudaMemcpyHostToDevice));
  const int rc = pattern_run(device, (const uint8_t*)lines.p, h_lines, n_blocks, line_size, cache_blocks, (uint16_t*)sizes.p, out, kernel_ms);
  if (rc == MPC_OK && h_sizes && n_blocks) PAT_CUDA(cudaMemcpy(h_sizes, sizes.p, n_blocks * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  return rc;
}
