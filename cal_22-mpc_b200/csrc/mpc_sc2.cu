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

#include <algorithm>
#include <cstring>
#include <cub/cub.cuh>
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

__global__ void __launch_bounds__(kThreads)
sc2_lookup_kernel(const uint4* __restrict__ lines, uint64_t n_blocks, uint64_t sampling, const uint32_t* __restrict__ g_syms,
                  const uint8_t* __restrict__ g_lens, int k, uint16_t* __restrict__ sizes, unsigned long long* __restrict__ total) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* s_stage = reinterpret_cast<uint4*>(smem_raw);
  __shared__ uint32_t s_syms[1024];
  __shared__ uint8_t s_lens[1024];
  for (int i = threadIdx.x; i < 1024; i += kThreads) {
    s_syms[i] = i < k ? g_syms[i] : 0xffffffffu;
    s_lens[i] = i < k ? g_lens[i] : 33;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long bits = 0;
  tile::for_each_block(lines, n_blocks, s_stage + warp * tile::kStages * 256, kWarps,
                       [&](const uint32_t (&x)[32], uint64_t blk, bool valid) {
    uint32_t size = 0;
    if (blk < sampling) {
      size = 33u * 32u;  // sampling phase: every word is a miss (SC2.cpp:315-323 with an empty code map)
    } else {
#pragma unroll 4
      for (int j = 0; j < 32; j++) {
        const uint32_t v = x[j];
        int lo = 0, hi = k - 1;
        uint32_t len = 33;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          const uint32_t s = s_syms[mid];
          if (s == v) { len = s_lens[mid]; break; }
          if (s < v) lo = mid + 1; else hi = mid - 1;
        }
        size += len;
      }
    }
    if (valid) {
      if (sizes) sizes[blk] = (uint16_t)size;
      bits += size;
    }
  });
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

extern "C" int mpc_sc2_run_device(int device, const uint8_t* d_lines, uint64_t n_blocks, uint32_t line_size, uint64_t sampling_lines,
                                  uint16_t* d_sizes, mpc_variant_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (!out || (n_blocks && !d_lines)) return fail(MPC_E_ARG, "null argument");
  if (line_size != 128) return fail(MPC_E_ARG, "the GPU variants are built for 128-byte blocks");
  if ((uintptr_t)d_lines & 15) return fail(MPC_E_ARG, "lines must be 16-byte aligned");
  SC2_CUDA(cudaSetDevice(device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaEvent_t e0, e1;
  SC2_CUDA(cudaEventCreate(&e0));
  SC2_CUDA(cudaEventCreate(&e1));
  SC2_CUDA(cudaEventRecord(e0, 0));
  std::vector<uint32_t> syms;
  std::vector<uint8_t> lens;
  if (n_blocks > sampling_lines && sampling_lines > 0) {
    // ---- phase 1: histogram of the sampled words = sort + run-length encode ----
    const uint64_t nw64 = sampling_lines * 32ull;
    if (nw64 > 0x7fffffffull) return fail(MPC_E_ARG, "sampling window too large");
    const int nw = (int)nw64;
    uint32_t *d_sorted = nullptr, *d_unique = nullptr, *d_counts = nullptr;
    int* d_runs = nullptr;
    uint64_t *d_keys = nullptr, *d_keys_sorted = nullptr;
    void* d_temp = nullptr;
    size_t temp1 = 0, temp2 = 0, temp3 = 0;
    SC2_CUDA(cudaMalloc(&d_sorted, (size_t)nw * 4));
    SC2_CUDA(cudaMalloc(&d_unique, (size_t)nw * 4));
    SC2_CUDA(cudaMalloc(&d_counts, (size_t)nw * 4));
    SC2_CUDA(cudaMalloc(&d_runs, sizeof(int)));
    const uint32_t* d_words = reinterpret_cast<const uint32_t*>(d_lines);
    cub::DeviceRadixSort::SortKeys(nullptr, temp1, d_words, d_sorted, nw);
    cub::DeviceRunLengthEncode::Encode(nullptr, temp2, d_sorted, d_unique, d_counts, d_runs, nw);
    cub::DeviceRadixSort::SortKeys(nullptr, temp3, d_keys, d_keys_sorted, nw);
    const size_t temp = std::max(temp1, std::max(temp2, temp3));
    SC2_CUDA(cudaMalloc(&d_temp, temp));
    size_t t = temp;
    SC2_CUDA(cub::DeviceRadixSort::SortKeys(d_temp, t, d_words, d_sorted, nw));
    t = temp;
    SC2_CUDA(cub::DeviceRunLengthEncode::Encode(d_temp, t, d_sorted, d_unique, d_counts, d_runs, nw));
    int runs = 0;
    SC2_CUDA(cudaMemcpy(&runs, d_runs, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<std::pair<uint32_t, uint64_t>> sf;
    if (runs > 1024) {
      // keep the 1024 largest by (count, symbol): the reference erases in ascending (freq, symbol) order (SC2.cpp:294-307)
      SC2_CUDA(cudaMalloc(&d_keys, (size_t)runs * 8));
      SC2_CUDA(cudaMalloc(&d_keys_sorted, (size_t)runs * 8));
      pack_count_symbol<<<(runs + 255) / 256, 256>>>(d_unique, d_counts, d_keys, (uint32_t)runs);
      t = temp;
      SC2_CUDA(cub::DeviceRadixSort::SortKeys(d_temp, t, d_keys, d_keys_sorted, runs));
      std::vector<uint64_t> top(1024);
      SC2_CUDA(cudaMemcpy(top.data(), d_keys_sorted + (runs - 1024), 1024 * 8, cudaMemcpyDeviceToHost));
      for (uint64_t k : top) sf.emplace_back((uint32_t)k, k >> 32);
      cudaFree(d_keys);
      cudaFree(d_keys_sorted);
    } else {
      std::vector<uint32_t> u(runs), c(runs);
      SC2_CUDA(cudaMemcpy(u.data(), d_unique, (size_t)runs * 4, cudaMemcpyDeviceToHost));
      SC2_CUDA(cudaMemcpy(c.data(), d_counts, (size_t)runs * 4, cudaMemcpyDeviceToHost));
      for (int i = 0; i < runs; i++) sf.emplace_back(u[i], c[i]);
    }
    cudaFree(d_temp);
    cudaFree(d_sorted);
    cudaFree(d_unique);
    cudaFree(d_counts);
    cudaFree(d_runs);
    std::sort(sf.begin(), sf.end());  // ascending symbol = std::map iteration order
    sc2_code_lengths(sf, &syms, &lens);
  }
  // ---- phase 2: per-block lookup ----
  uint32_t* d_syms = nullptr;
  uint8_t* d_lens = nullptr;
  unsigned long long* d_total = nullptr;
  SC2_CUDA(cudaMalloc(&d_syms, 1024 * 4));
  SC2_CUDA(cudaMalloc(&d_lens, 1024));
  SC2_CUDA(cudaMalloc(&d_total, 8));
  SC2_CUDA(cudaMemset(d_total, 0, 8));
  if (!syms.empty()) {
    SC2_CUDA(cudaMemcpy(d_syms, syms.data(), syms.size() * 4, cudaMemcpyHostToDevice));
    SC2_CUDA(cudaMemcpy(d_lens, lens.data(), lens.size(), cudaMemcpyHostToDevice));
  }
  if (n_blocks) {
    const size_t smem = (size_t)kWarps * tile::kStages * tile::kTileBytes;
    SC2_CUDA(cudaFuncSetAttribute(sc2_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SC2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sc2_lookup_kernel, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t tiles = (n_blocks + 31) / 32;
    uint64_t grid = std::min<uint64_t>((uint64_t)sms * per_sm, (tiles + kWarps - 1) / kWarps);
    // lines [0, S) are sampling lines; when the dump has no more than S lines the tree is never built (SC2.cpp:285-313)
    const uint64_t s_eff = (n_blocks > sampling_lines) ? sampling_lines : n_blocks;
    sc2_lookup_kernel<<<(unsigned)grid, kThreads, smem>>>(reinterpret_cast<const uint4*>(d_lines), n_blocks, s_eff, d_syms, d_lens,
                                                          (int)syms.size(), d_sizes, d_total);
    SC2_CUDA(cudaGetLastError());
  }
  SC2_CUDA(cudaEventRecord(e1, 0));
  unsigned long long total = 0;
  SC2_CUDA(cudaMemcpy(&total, d_total, 8, cudaMemcpyDeviceToHost));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_syms);
  cudaFree(d_lens);
  cudaFree(d_total);
  memset(out, 0, sizeof(*out));
  out->blocks = n_blocks;
  out->original_bits = n_blocks * 8ull * line_size;
  out->compressed_bits = total;
  out->counts[0] = syms.size();  // symbols that received a code
  if (kernel_ms) *kernel_ms = ms;
  return MPC_OK;
}

extern "C" int mpc_sc2_run_host(int device, const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size, uint64_t sampling_lines,
                                uint16_t* h_sizes, mpc_variant_stats* out, float* kernel_ms) {
  using namespace mpc;
  if (n_blocks && !h_lines) return fail(MPC_E_ARG, "null lines");
  SC2_CUDA(cudaSetDevice(device));
  uint8_t* d_lines = nullptr;
  uint16_t* d_sizes = nullptr;
  const size_t bytes = (size_t)n_blocks * line_size;
  SC2_CUDA(cudaMalloc(&d_lines, bytes ? bytes : 16));
  if (h_sizes) SC2_CUDA(cudaMalloc(&d_sizes, (n_blocks ? n_blocks : 1) * sizeof(uint16_t)));
  if (bytes) SC2_CUDA(cudaMemcpy(d_lines, h_lines, bytes, cudaMemcpyHostToDevice));
  int rc = mpc_sc2_run_device(device, d_lines, n_blocks, line_size, sampling_lines, d_sizes, out, kernel_ms);
  if (rc == MPC_OK && h_sizes && n_blocks) cudaMemcpy(h_sizes, d_sizes, n_blocks * sizeof(uint16_t), cudaMemcpyDeviceToHost);
  cudaFree(d_lines);
  if (d_sizes) cudaFree(d_sizes);
  return rc;
}

// CPACK: host, sequential (see the file header).  counts = ZZZZ, XXXX, MMMM, MMXX, ZZZX, MMMX (CPACK.h:119-127).
extern "C" int mpc_cpack_run_host(const uint8_t* h_lines, uint64_t n_blocks, uint32_t line_size, uint16_t* h_sizes,
                                  mpc_variant_stats* out) {
  if (!out || (n_blocks && !h_lines) || line_size % 4) return MPC_E_ARG;
  static const uint32_t kLen[6] = {2, 34, 6, 24, 12, 16};
  uint32_t dict[16] = {0};  // FIFO, oldest at `head`; words kept as little-endian 32-bit values
  int head = 0;
  memset(out, 0, sizeof(*out));
  const uint32_t W = line_size / 4;
  for (uint64_t b = 0; b < n_blocks; b++) {
    uint32_t size = 0;
    for (uint32_t i = 0; i < W; i++) {
      uint32_t w;
      memcpy(&w, h_lines + b * line_size + 4 * i, 4);
      int pat = -1;
      if ((w & 0x00ffffffu) == 0) {
        pat = (w == 0) ? 0 : 4;  // zzzz / zzzx
      } else {
        for (int j = 0; j < 16 && pat < 0; j++) {
          const uint32_t d = dict[(head + j) & 15];
          if (((w ^ d) & 0x0000ffffu) == 0) pat = ((w ^ d) & 0x00ff0000u) ? 3 : (((w ^ d) & 0xff000000u) ? 5 : 2);
        }
        if (pat < 0) {
          pat = 1;
          dict[head] = w;
          head = (head + 1) & 15;
        }
      }
      size += kLen[pat];
      out->counts[pat]++;
    }
    if (h_sizes) h_sizes[b] = (uint16_t)size;
    out->compressed_bits += size;
  }
  out->blocks = n_blocks;
  out->original_bits = n_blocks * 8ull * line_size;
  return MPC_OK;
}
