// SC2 (Huffman over 32-bit words) and CPACK size models behind the C ABI.
//
// SC2 (reference src/compressor/SC2.cpp:270-334) is two-phase: the first S lines only feed a symbol histogram and
// cost 33 bits per word; at line S the <= 1024 most frequent symbols get Huffman code lengths; from then on a word
// costs its code length or 33 bits.  Here: phase 1 = device radix sort + run-length encode of the sampled words
// (CUB) and a second sort by (count, symbol) to keep the 1024 survivors; the tree is built on the host with the
// reference's own array min-heap rules (SC2.cpp:24-126: strict <, left child before right, parent = ceil(i/2)-1),
// because tie-breaking inside the heap decides the code lengths; phase 2 = one thread per block, binary search in a
// shared-memory copy of the table.
//
// CPACK (CPACK.cpp:7-101) carries a 16-entry FIFO dictionary across lines, so its result depends on the order of
// all earlier words: it is sequential by construction and runs on the host (mpc_cpack_run_host).
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>  // host side of CPACK: the sixteen dictionary entries are searched with two compares
#endif

#include <algorithm>
#include <cstring>
#include <cub/cub.cuh>
#include <mutex>
#include <string>
#include <vector>

#include "mpc_capi.h"
#include "mpc_tile.cuh"

namespace mpc {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

thread_local std::string g_err;
int fail(int code, const std::string& what) { g_err = what; return code; }
#define SC2_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) return fail(MPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

__global__ void pack_count_symbol(const uint32_t* __restrict__ sym, const uint32_t* __restrict__ cnt, uint64_t* __restrict__ key,
                                  uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = ((uint64_t)cnt[i] << 32) | sym[i];
}

// Code table for the lookup kernel: open addressing, 4096 slots for <= 1024 symbols (load <= 25 %, ~1.2 probes per word),
// built on the host from the (symbol, length) list.  s_len holds length + 1 (a single-symbol tree has length 0), 0 = empty.
constexpr int kHashSlots = 4096;
__host__ __device__ __forceinline__ uint32_t sc2_hash(uint32_t v) { return (v * 0x9E3779B1u) >> 20; }

// W = words per line: a thread holds 128 bytes = 32 / W consecutive lines (SC2 works word by word, SC2.cpp:315-330, so any line size)
template <int W>
__global__ void __launch_bounds__(kThreads)
sc2_lookup_kernel(const uint4* __restrict__ lines, uint64_t n_blocks, uint64_t first_block, uint64_t sampling, const uint32_t* __restrict__ g_keys,
                  const uint8_t* __restrict__ g_lens, uint16_t* __restrict__ sizes, unsigned long long* __restrict__ total) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* s_stage = reinterpret_cast<uint4*>(smem_raw);
  __shared__ uint2 s_tab[kHashSlots];  // (symbol, code length + 1), length field 0 = empty slot: one 8-byte load per probe
  for (int i = threadIdx.x; i < kHashSlots; i += kThreads) s_tab[i] = make_uint2(g_keys[i], (uint32_t)g_lens[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long bits = 0;
  constexpr int S = 32 / W;
  tile::for_each_block(lines, (n_blocks + S - 1) / S, s_stage + warp * tile::kStages * 256, kWarps,
                       [&](const uint32_t (&x128)[32], uint64_t unit, bool) {
#pragma unroll
   for (int sub = 0; sub < S; sub++) {
    uint32_t x[32];
#pragma unroll
    for (int i = 0; i < W; i++) x[i] = x128[sub * W + i];
    const uint64_t blk = unit * S + sub;
    const bool valid = blk < n_blocks;
    uint32_t size = 0;
    if (first_block + blk < sampling) {
      size = 33u * (uint32_t)W;  // sampling phase: every word is a miss (SC2.cpp:315-323 with an empty code map)
    } else {
      // first probe of all 32 words with no branch in between (32 independent loads in flight); only a word whose first slot
      // holds another symbol walks on
      uint32_t pending = 0;  // bit j: word j has to look at further slots
#pragma unroll
      for (int j = 0; j < W; j++) {
        const uint2 e = s_tab[sc2_hash(x[j])];
        const bool hit = e.y != 0u && e.x == x[j];
        size += hit ? e.y - 1u : 33u;
        pending |= (e.y != 0u && !hit) ? (1u << j) : 0u;
      }
      if (pending) {
#pragma unroll
        for (int j = 0; j < W; j++) {
          if ((pending >> j) & 1u) {
            const uint32_t v = x[j];
            uint32_t idx = (sc2_hash(v) + 1u) & (kHashSlots - 1);
            while (true) {
              const uint2 e = s_tab[idx];
              if (e.y == 0u) break;                                    // empty slot: not in the table (33 already counted)
              if (e.x == v) { size += e.y - 1u - 33u; break; }           // found after all: replace the 33
              idx = (idx + 1u) & (kHashSlots - 1);
            }
          }
        }
      }
    }
    if (valid) {
      if (sizes) sizes[blk] = (uint16_t)size;
      bits += size;
    }
   }  // sub-lines
  }, n_blocks * (uint64_t)(W / 4));
  for (int o = 16; o; o >>= 1) bits += __shfl_down_sync(0xffffffffu, bits, o);
  if (lane == 0 && bits) atomicAdd(total, bits);
}

// huffman::MinHeap + BuildHuffmanTree + GetHuffmanCode (SC2.cpp:24-162) on (symbol, freq) pairs given in ascending
// symbol order (the iteration order of the reference's std::map).  Returns (symbol, code length) sorted by symbol.
struct Node { int64_t symbol; uint64_t freq; int left, right; };

void sc2_code_lengths(const std::vector<std::pair<uint32_t, uint64_t>>& sf, std::vector<uint32_t>* syms, std::vector<uint8_t>* lens) {
  syms->clear();
  lens->clear();
  if (sf.empty()) return;
  std::vector<Node> pool;
  pool.reserve(2 * sf.size() + 1);
  std::vector<int> heap;
  for (auto& p : sf) { pool.push_back({(int64_t)p.first, p.second, -1, -1}); heap.push_back((int)pool.size() - 1); }
  int size = (int)heap.size();
  auto freq = [&](int i) { return pool[heap[i]].freq; };
  auto sift_down = [&](int index) {  // minHeapify
    for (;;) {
      int m = index, l = 2 * index + 1, r = 2 * index + 2;
      if (l <= size - 1 && freq(l) < freq(m)) m = l;
      if (r <= size - 1 && freq(r) < freq(m)) m = r;
      if (m == index) return;
      std::swap(heap[index], heap[m]);
      index = m;
    }
  };
  for (int i = size / 2 - 1; i >= 0; i--) sift_down(i);
  auto extract_min = [&]() {
    int top = heap[0];
    std::swap(heap[0], heap[size - 1]);
    size--;
    sift_down(0);
    return top;
  };
  while (size > 1) {
    const int l = extract_min(), r = extract_min();
    pool.push_back({-1, pool[l].freq + pool[r].freq, l, r});
    heap[size++] = (int)pool.size() - 1;
    for (int i = size - 1; i > 0;) {
      const int p = (i + 1) / 2 - 1;
      if (!(freq(p) > freq(i))) break;
      std::swap(heap[i], heap[p]);
      i = p;
    }
  }
  std::vector<std::pair<uint32_t, uint8_t>> out;
  std::vector<std::pair<int, int>> stack{{heap[0], 0}};
  while (!stack.empty()) {
    auto [nd, depth] = stack.back();
    stack.pop_back();
    if (pool[nd].left < 0 && pool[nd].right < 0) { out.emplace_back((uint32_t)pool[nd].symbol, (uint8_t)depth); continue; }
    stack.push_back({pool[nd].right, depth + 1});
    stack.push_back({pool[nd].left, depth + 1});
  }
  std::sort(out.begin(), out.end());
  for (auto& p : out) { syms->push_back(p.first); lens->push_back(p.second); }
}

}  // namespace
}  // namespace mpc

extern "C" const char* mpc_sc2_error(void) { return mpc::g_err.c_str(); }

namespace mpc {
namespace {

// Per-device workspace, kept for the life of the process: the sort buffers of the sampling phase (grown on demand), the code
// table, the counters and the two timing events -- no allocation or event creation per call.
struct Sc2Workspace {
  int device = -1;
  size_t words = 0, temp_bytes = 0, keys = 0;
  uint32_t *d_sorted = nullptr, *d_unique = nullptr, *d_counts = nullptr;
  int* d_runs = nullptr;
  uint64_t *d_keys = nullptr, *d_keys_sorted = nullptr;
  void* d_temp = nullptr;
  uint32_t* d_tab_keys = nullptr;
  uint8_t* d_tab_lens = nullptr;
  unsigned long long* d_total = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};
std::mutex g_ws_mutex;
Sc2Workspace g_ws[16];

int ws_get(int device, Sc2Workspace** out) {
  if (device < 0 || device >= 16) return fail(MPC_E_ARG, "device index out of range");
  Sc2Workspace& w = g_ws[device];
  if (w.device != device) {
    // all or nothing: a failure half way leaves no buffer behind for the next call to leak or to trust
    cudaError_t e = cudaMalloc(&w.d_runs, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&w.d_tab_keys, kHashSlots * 4);
    if (e == cudaSuccess) e = cudaMalloc(&w.d_tab_lens, kHashSlots);
    if (e == cudaSuccess) e = cudaMalloc(&w.d_total, 8);
    if (e == cudaSuccess) e = cudaEventCreate(&w.e0);
    if (e == cudaSuccess) e = cudaEventCreate(&w.e1);
    if (e != cudaSuccess) {
      cudaFree(w.d_runs); cudaFree(w.d_tab_keys); cudaFree(w.d_tab_lens); cudaFree(w.d_total);
      if (w.e0) cudaEventDestroy(w.e0);
      if (w.e1) cudaEventDestroy(w.e1);
      w.d_runs = nullptr; w.d_tab_keys = nullptr; w.d_tab_lens = nullptr; w.d_total = nullptr; w.e0 = w.e1 = nullptr;
      return fail(MPC_E_CUDA, std::string("SC2 workspace: ") + cudaGetErrorString(e));
    }
    w.device = device;
  }
  *out = &w;
  return MPC_OK;
}

int ws_reserve(Sc2Workspace& w, size_t words) {
  if (words <= w.words) return MPC_OK;
  if (w.d_sorted) { cudaFree(w.d_sorted); cudaFree(w.d_unique); cudaFree(w.d_counts); cudaFree(w.d_temp); }
  if (w.d_keys) { cudaFree(w.d_keys); cudaFree(w.d_keys_sorted); }
  w.d_sorted = w.d_unique = w.d_counts = nullptr;
  w.d_keys = w.d_keys_sorted = nullptr;
  w.d_temp = nullptr;
  w.words = 0;
  SC2_CUDA(cudaMalloc(&w.d_sorted, words * 4));
  SC2_CUDA(cudaMalloc(&w.d_unique, words * 4));
  SC2_CUDA(cudaMalloc(&w.d_counts, words * 4));
  SC2_CUDA(cudaMalloc(&w.d_keys, words * 8));
  SC2_CUDA(cudaMalloc(&w.d_keys_sorted, words * 8));
  size_t t1 = 0, t2 = 0, t3 = 0;
  const int nw = (int)words;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, w.d_sorted, w.d_sorted, nw);
  cub::DeviceRunLengthEncode::Encode(nullptr, t2, w.d_sorted, w.d_unique, w.d_counts, w.d_runs, nw);
  cub::DeviceRadixSort::SortKeys(nullptr, t3, w.d_keys, w.d_keys_sorted, nw);
  w.temp_bytes = std::max(t1, std::max(t2, t3));
  SC2_CUDA(cudaMalloc(&w.d_temp, w.temp_bytes));
  w.words = words;
  return MPC_OK;
}

}  // namespace
}  // namespace mpc

// Phase 1 (SC2.cpp:285-313): the code table from the first sampling_lines lines, which must be resident at d_lines.
extern "C" int mpc_sc2_build_table(int device, const uint8_t* d_lines, uint64_t sampling_lines, uint32_t line_size, mpc_sc2_table* table) {
  using namespace mpc;
  if (!table || (sampling_lines && !d_lines)) return fail(MPC_E_ARG, "null argument");
  if (line_size != 32 && line_size != 64 && line_size != 128) return fail(MPC_E_ARG, "line size must be 32, 64 or 128 bytes");
  memset(table, 0, sizeof(*table));
  if (sampling_lines == 0) return MPC_OK;
  const uint64_t nw64 = sampling_lines * (uint64_t)(line_size / 4);
  if (nw64 > 0x7fffffffull) return fail(MPC_E_ARG, "sampling window too large");
  const int nw = (int)nw64;
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  SC2_CUDA(cudaSetDevice(device));
  Sc2Workspace* wp = nullptr;
  int rc = ws_get(device, &wp);
  if (rc != MPC_OK) return rc;
  Sc2Workspace& w = *wp;
  if ((rc = ws_reserve(w, (size_t)nw)) != MPC_OK) return rc;
  // histogram of the sampled words = sort + run-length encode
  const uint32_t* d_words = reinterpret_cast<const uint32_t*>(d_lines);
  size_t t = w.temp_bytes;
  SC2_CUDA(cub::DeviceRadixSort::SortKeys(w.d_temp, t, d_words, w.d_sorted, nw));
  t = w.temp_bytes;
  SC2_CUDA(cub::DeviceRunLengthEncode::Encode(w.d_temp, t, w.d_sorted, w.d_unique, w.d_counts, w.d_runs, nw));
  int runs = 0;
  SC2_CUDA(cudaMemcpy(&runs, w.d_runs, sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<std::pair<uint32_t, uint64_t>> sf;
  if (runs > 1024) {
    // keep the 1024 largest by (count, symbol): the reference erases in ascending (freq, symbol) order (SC2.cpp:294-307)
    pack_count_symbol<<<(runs + 255) / 256, 256>>>(w.d_unique, w.d_counts, w.d_keys, (uint32_t)runs);
    t = w.temp_bytes;
    SC2_CUDA(cub::DeviceRadixSort::SortKeys(w.d_temp, t, w.d_keys, w.d_keys_sorted, runs));
    std::vector<uint64_t> top(1024);
    SC2_CUDA(cudaMemcpy(top.data(), w.d_keys_sorted + (runs - 1024), 1024 * 8, cudaMemcpyDeviceToHost));
    for (uint64_t k : top) sf.emplace_back((uint32_t)k, k >> 32);
  } else {
    std::vector<uint32_t> u((size_t)runs), c((size_t)runs);
    SC2_CUDA(cudaMemcpy(u.data(), w.d_unique, (size_t)runs * 4, cudaMemcpyDeviceToHost));
    SC2_CUDA(cudaMemcpy(c.data(), w.d_counts, (size_t)runs * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < runs; i++) sf.emplace_back(u[(size_t)i], c[(size_t)i]);
  }
  std::sort(sf.begin(), sf.end());  // ascending symbol = std::map iteration order
  std::vector<uint32_t> syms;
  std::vector<uint8_t> lens;
  sc2_code_lengths(sf, &syms, &lens);
  table->n = (uint32_t)syms.size();
  for (size_t i = 0; i < syms.size(); i++) { table->symbols[i] = syms[i]; table->lengths[i] = lens[i]; }
  return MPC_OK;
}

// Phase 2 (SC2.cpp:315-330): n_blocks lines whose first one is line `first_block` of the dump (a shard of it, or all of it):
// lines below sampling_lines cost 33 bits per word, the others the code length of each word or 33.
extern "C" int mpc_sc2_apply_device(int device, const uint8_t* d_lines, uint64_t n_blocks, uint64_t first_block, uint64_t sampling_lines,
                                    uint32_t line_size, const mpc_sc2_table* table, uint16_t* d_sizes, mpc_variant_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (!out || !table || (n_blocks && !d_lines)) return fail(MPC_E_ARG, "null argument");
  if (line_size != 32 && line_size != 64 && line_size != 128) return fail(MPC_E_ARG, "line size must be 32, 64 or 128 bytes");
  if ((uintptr_t)d_lines & 15) return fail(MPC_E_ARG, "lines must be 16-byte aligned");
  if (table->n > 1024) return fail(MPC_E_ARG, "code table holds more than 1024 symbols");
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  SC2_CUDA(cudaSetDevice(device));
  Sc2Workspace* wp = nullptr;
  int rc = ws_get(device, &wp);
  if (rc != MPC_OK) return rc;
  Sc2Workspace& w = *wp;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  std::vector<uint32_t> keys(kHashSlots, 0u);
  std::vector<uint8_t> lens(kHashSlots, 0);
  for (uint32_t i = 0; i < table->n; i++) {
    uint32_t idx = sc2_hash(table->symbols[i]);
    while (lens[idx]) idx = (idx + 1u) & (kHashSlots - 1);
    keys[idx] = table->symbols[i];
    lens[idx] = (uint8_t)(table->lengths[i] + 1);
  }
  SC2_CUDA(cudaEventRecord(w.e0, 0));
  SC2_CUDA(cudaMemcpyAsync(w.d_tab_keys, keys.data(), kHashSlots * 4, cudaMemcpyHostToDevice, 0));
  SC2_CUDA(cudaMemcpyAsync(w.d_tab_lens, lens.data(), kHashSlots, cudaMemcpyHostToDevice, 0));
  SC2_CUDA(cudaMemsetAsync(w.d_total, 0, 8, 0));
  if (n_blocks) {
    const size_t smem = (size_t)kWarps * tile::kStages * tile::kTileBytes;
    auto launch = [&](auto kernel, uint64_t lines_per_unit) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      int per_sm = 1;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
      if (e != cudaSuccess) return e;
      if (per_sm < 1) per_sm = 1;
      const uint64_t units = (n_blocks + lines_per_unit - 1) / lines_per_unit;  // 128-byte units
      const uint64_t tiles = (units + 31) / 32;
      const uint64_t grid = std::min<uint64_t>((uint64_t)sms * per_sm, (tiles + kWarps - 1) / kWarps);
      kernel<<<(unsigned)grid, kThreads, smem>>>(reinterpret_cast<const uint4*>(d_lines), n_blocks, first_block, sampling_lines,
                                                 w.d_tab_keys, w.d_tab_lens, d_sizes, w.d_total);
      return cudaGetLastError();
    };
    if (line_size == 32) SC2_CUDA(launch(sc2_lookup_kernel<8>, 4));
    else if (line_size == 64) SC2_CUDA(launch(sc2_lookup_kernel<16>, 2));
    else SC2_CUDA(launch(sc2_lookup_kernel<32>, 1));
  }
  SC2_CUDA(cudaEventRecord(w.e1, 0));
  unsigned long long total = 0;
  SC2_CUDA(cudaMemcpy(&total, w.d_total, 8, cudaMemcpyDeviceToHost));
  float ms = 0;
  SC2_CUDA(cudaEventElapsedTime(&ms, w.e0, w.e1));
  memset(out, 0, sizeof(*out));
  out->blocks = n_blocks;
  out->original_bits = n_blocks * 8ull * line_size;
  out->compressed_bits = total;
  out->counts[0] = table->n;  // symbols that received a code
  if (kernel_ms) *kernel_ms = ms;
  return MPC_OK;
}

extern "C" int mpc_sc2_run_device(int device, const uint8_t* d_lines, uint64_t n_blocks, uint32_t line_size, uint64_t sampling_lines,
                                  uint16_t* d_sizes, mpc_variant_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (!out || (n_blocks && !d_lines)) return fail(MPC_E_ARG, "null argument");
  mpc_sc2_table table;
  memset(&table, 0, sizeof(table));
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // the whole call on the device's clock: table build (sorts, host tree) + lookup
  SC2_CUDA(cudaSetDevice(device));
  SC2_CUDA(cudaEventCreateWithFlags(&t0, cudaEventDefault));
  SC2_CUDA(cudaEventCreateWithFlags(&t1, cudaEventDefault));
  cudaEventRecord(t0, 0);
  int rc = MPC_OK;
  // when the dump has no more than S lines the tree is never built (SC2.cpp:285-313): every line is a sampling line
  if (n_blocks > sampling_lines && sampling_lines > 0) rc = mpc_sc2_build_table(device, d_lines, sampling_lines, line_size, &table);
  float ms_apply = 0;
  if (rc == MPC_OK) rc = mpc_sc2_apply_device(device, d_lines, n_blocks, 0, sampling_lines, line_size, &table, d_sizes, out, &ms_apply);
  cudaEventRecord(t1, 0);
  cudaEventSynchronize(t1);
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  if (rc == MPC_OK && kernel_ms) *kernel_ms = ms;
  return rc;
}

// Host dump: the sampling window (at most 10^6 lines = 128 MB) goes to the device first for the table, then the dump streams
// through two device buffers in chunks -- dumps larger than HBM work, every copy's return code is checked.
extern "C" int mpc_sc2_run_host(int device, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size, uint64_t sampling_lines,
                                uint16_t* h_sizes, mpc_variant_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (!out || (n_blocks && !h_lines)) return fail(MPC_E_ARG, "null argument");
  if (line_size != 32 && line_size != 64 && line_size != 128) return fail(MPC_E_ARG, "line size must be 32, 64 or 128 bytes");
  SC2_CUDA(cudaSetDevice(device));
  const uint64_t chunk = (256ull << 20) / line_size;
  const bool build = n_blocks > sampling_lines && sampling_lines > 0;
  const uint64_t cap = std::max<uint64_t>(std::min<uint64_t>(chunk, n_blocks ? n_blocks : 1), build ? sampling_lines : 1);
  uint8_t* d_lines = nullptr;
  uint16_t* d_sizes = nullptr;
  SC2_CUDA(cudaMalloc(&d_lines, (size_t)cap * line_size));
  if (h_sizes && cudaMalloc(&d_sizes, (size_t)cap * sizeof(uint16_t)) != cudaSuccess) { cudaFree(d_lines); return fail(MPC_E_CUDA, "cudaMalloc (sizes)"); }
  auto done = [&](int rc) { cudaFree(d_lines); if (d_sizes) cudaFree(d_sizes); return rc; };
  mpc_sc2_table table;
  memset(&table, 0, sizeof(table));
  float ms_total = 0;
  if (build) {
    if (cudaMemcpy(d_lines, h_lines, (size_t)sampling_lines * line_size, cudaMemcpyHostToDevice) != cudaSuccess) return done(fail(MPC_E_CUDA, "H2D copy of the sampling window"));
    const int rc = mpc_sc2_build_table(device, d_lines, sampling_lines, line_size, &table);
    if (rc != MPC_OK) return done(rc);
  }
  memset(out, 0, sizeof(*out));
  for (uint64_t lo = 0; lo < n_blocks || lo == 0; lo += chunk) {
    const uint64_t nb = std::min<uint64_t>(chunk, n_blocks - lo);
    if (nb && cudaMemcpy(d_lines, h_lines + lo * line_size, (size_t)nb * line_size, cudaMemcpyHostToDevice) != cudaSuccess) return done(fail(MPC_E_CUDA, "H2D copy"));
    mpc_variant_stats part;
    float ms = 0;
    const int rc = mpc_sc2_apply_device(device, d_lines, nb, lo, sampling_lines, line_size, &table, d_sizes, &part, &ms);
    if (rc != MPC_OK) return done(rc);
    if (h_sizes && nb && cudaMemcpy(h_sizes + lo, d_sizes, (size_t)nb * sizeof(uint16_t), cudaMemcpyDeviceToHost) != cudaSuccess) return done(fail(MPC_E_CUDA, "D2H copy"));
    out->blocks += part.blocks;
    out->original_bits += part.original_bits;
    out->compressed_bits += part.compressed_bits;
    out->counts[0] = part.counts[0];
    ms_total += ms;
    if (n_blocks == 0) break;
  }
  if (kernel_ms) *kernel_ms = ms_total;
  return done(MPC_OK);
}

// CPACK: host, sequential (see the file header).  counts = ZZZZ, XXXX, MMMM, MMXX, ZZZX, MMMX (CPACK.h:119-127).
extern "C" int mpc_cpack_run_host(const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size, uint16_t* h_sizes,
                                  mpc_variant_stats* out) {
  if (!out || (n_blocks && !h_lines) || line_size % 4) return MPC_E_ARG;
  static const uint32_t kLen[6] = {2, 34, 6, 24, 12, 16};
  uint32_t dict[16] = {0};  // FIFO, oldest at `head`; words kept as little-endian 32-bit values
  alignas(16) uint16_t low[16] = {0};  // their low halfwords: what a dictionary match is decided on (bytes 0-1, CPACK.cpp:45-70)
  int head = 0;
  memset(out, 0, sizeof(*out));
  const uint32_t W = line_size / 4;
  uint64_t counts[6] = {0, 0, 0, 0, 0, 0};
  for (uint64_t b = 0; b < n_blocks; b++) {
    uint32_t size = 0;
    for (uint32_t i = 0; i < W; i++) {
      uint32_t w;
      memcpy(&w, h_lines + b * line_size + 4 * i, 4);
      int pat = -1;
      if ((w & 0x00ffffffu) == 0) {
        pat = (w == 0) ? 0 : 4;  // zzzz / zzzx
      } else {
        // the first entry, oldest first, whose bytes 0-1 equal the word's
#if defined(__SSE2__)
        // all sixteen entries at once: compare the halfwords, one bit per entry, rotated so that bit 0 is the oldest entry
        const __m128i key = _mm_set1_epi16((short)(w & 0xffffu));
        const __m128i e0 = _mm_cmpeq_epi16(_mm_load_si128(reinterpret_cast<const __m128i*>(low)), key);
        const __m128i e1 = _mm_cmpeq_epi16(_mm_load_si128(reinterpret_cast<const __m128i*>(low + 8)), key);
        const uint32_t m = (uint32_t)_mm_movemask_epi8(_mm_packs_epi16(e0, e1));  // bit j = entry j matches
        const uint32_t r = ((m >> head) | (m << (16 - head))) & 0xffffu;
        if (r) {
          const uint32_t x = w ^ dict[(head + __builtin_ctz(r)) & 15];
          pat = (x & 0x00ff0000u) ? 3 : ((x & 0xff000000u) ? 5 : 2);
        }
#else
        for (int j = 0; j < 16 && pat < 0; j++) {
          const uint32_t d = dict[(head + j) & 15];
          if (((w ^ d) & 0x0000ffffu) == 0) pat = ((w ^ d) & 0x00ff0000u) ? 3 : (((w ^ d) & 0xff000000u) ? 5 : 2);
        }
#endif
        if (pat < 0) {
          pat = 1;
          dict[head] = w;
          low[head] = (uint16_t)(w & 0xffffu);
          head = (head + 1) & 15;
        }
      }
      size += kLen[pat];
      counts[pat]++;
    }
    if (h_sizes) h_sizes[b] = (uint16_t)size;
    out->compressed_bits += size;
  }
  for (int p = 0; p < 6; p++) out->counts[p] = counts[p];
  out->blocks = n_blocks;
  out->original_bits = n_blocks * 8ull * line_size;
  return MPC_OK;
}
