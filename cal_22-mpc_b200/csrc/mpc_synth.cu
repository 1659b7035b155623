// Synthetic memory-dump generator (device side).  Counter-based: every 128-byte block is a pure
// function of (kind, seed, global block index), so any shard of a 64 GB dump can be regenerated on
// the CPU for checking -- tools/gen_dump.py implements the same functions in numpy.
// Data classes follow SURVEY.md section 8d / BASELINE.md section 3 (zero pages, repeated word, smooth fp32,
// int32 index ramps, pointers, random bytes, sparse small ints, noisy fp32), in integer arithmetic only.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mpc_device.cuh"
#include "mpc_internal.h"

namespace mpc {
namespace {

__device__ __forceinline__ void synth_block(uint32_t* w, uint64_t b, uint64_t total, int kind, uint64_t seed) {
  const uint64_t key = mpcdev::splitmix64(seed ^ mpcdev::splitmix64(b));
  int c = kind;
  if (kind == MPC_SYN_MIXED_HASHED) c = (int)(key >> 61);
  else if (kind == MPC_SYN_MIXED_REGIONS) { c = (int)((b * 8ull) / (total ? total : 1)); if (c > 7) c = 7; }
  switch (c) {
    case MPC_SYN_ZERO:
      for (int k = 0; k < 32; k++) w[k] = 0;
      break;
    case MPC_SYN_WORDSAME: {
      uint32_t v = (uint32_t)(key >> 16);
      for (int k = 0; k < 32; k++) w[k] = v;
      break;
    }
    case MPC_SYN_SMOOTH_F32: {
      uint32_t v = 0x3F800000u + (uint32_t)((b * 2741ull) & 0x3FFFFFull);
      for (int k = 0; k < 32; k++) {
        uint64_t t = mpcdev::splitmix64(key + (uint64_t)k);
        v += (uint32_t)(t & 0x1FFFull) - 0x1000u;
        w[k] = v;
      }
      break;
    }
    case MPC_SYN_RAMP_I32: {
      uint32_t base = (uint32_t)(key & 0xFFFFFull);
      for (int k = 0; k < 32; k++) w[k] = base + 4u * (uint32_t)k;
      break;
    }
    case MPC_SYN_POINTER:
      for (int j = 0; j < 16; j++) {
        uint64_t t = mpcdev::splitmix64(key + (uint64_t)j);
        uint64_t v = 0x00007f0000000000ull + 8ull * (t & 0x3FFFFFFFull);
        w[2 * j] = (uint32_t)v;
        w[2 * j + 1] = (uint32_t)(v >> 32);
      }
      break;
    case MPC_SYN_RANDOM:
      for (int k = 0; k < 32; k++) w[k] = (uint32_t)mpcdev::splitmix64(key + (uint64_t)k);
      break;
    case MPC_SYN_SPARSE_I32:
      for (int k = 0; k < 32; k++) {
        uint64_t t = mpcdev::splitmix64(key + (uint64_t)k);
        w[k] = ((t >> 32) % 10ull < 6ull) ? 0u : (uint32_t)((int32_t)((t & 0xFFFFull) % 200ull) - 100);
      }
      break;
    default:  // MPC_SYN_NOISY_F32
      for (int k = 0; k < 32; k++) {
        uint64_t t = mpcdev::splitmix64(key + (uint64_t)k);
        uint32_t e = 120u + (uint32_t)((t >> 40) % 14ull);
        w[k] = ((uint32_t)(t >> 63) << 31) | (e << 23) | (uint32_t)(t & 0x7FFFFFull);
      }
      break;
  }
}

__global__ void mpc_synth_kernel(uint4* __restrict__ out, uint64_t first, uint64_t n, uint64_t total, int kind,
                                 uint64_t seed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t w[32];
    synth_block(w, first + i, total, kind, seed);
    uint4* dst = out + i * 8;
#pragma unroll
    for (int q = 0; q < 8; q++) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  }
}

}  // namespace

cudaError_t launch_synth(uint8_t* d_lines, uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks, int kind,
                         uint64_t seed, cudaStream_t stream) {
  if (n_blocks == 0) return cudaSuccess;
  uint64_t grid = (n_blocks + 127) / 128;
  if (grid > 148ull * 32) grid = 148ull * 32;
  mpc_synth_kernel<<<(unsigned)grid, 128, 0, stream>>>(reinterpret_cast<uint4*>(d_lines), first_block, n_blocks,
                                                       total_blocks, kind, seed);
  return cudaGetLastError();
}

}  // namespace mpc
