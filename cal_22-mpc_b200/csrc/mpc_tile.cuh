// Per-warp tile loader shared by the thread-per-block kernels: 32 consecutive 128-byte blocks (4 KiB) per tile,
// copied with coalesced 16-byte cp.async (LDGSTS) into a double-buffered, XOR-swizzled shared-memory stage and
// read back as one block per lane with 8 conflict-free LDS.128.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpc {
namespace tile {

constexpr int kTileBlocks = 32;
constexpr int kTileBytes = kTileBlocks * 128;
constexpr int kStages = 2;

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Calls body(x, blk, valid) for every block of the dump, one block per lane per tile; tiles are dealt round-robin
// to the warps of the grid.  `my_stage` = this warp's 2 x 4 KiB of shared memory.
// STAGES = 2: the next tile is requested before the current one is read (two 4 KiB stages per warp); STAGES = 1: it is requested right
// after the current tile sits in registers (the registers are the second buffer: half the shared memory, more CTAs per SM).
template <int STAGES = kStages, class Body>
// `valid_chunks` (16-byte chunks of the dump that exist; 0 = n_blocks * 8) bounds the reads when the last 128-byte unit is
// partial (lines of 32 / 64 bytes); the missing chunks arrive as zeros.
__device__ __forceinline__ void for_each_block(const uint4* __restrict__ lines, uint64_t n_blocks, uint4* my_stage,
                                               int warps_per_cta, Body body, uint64_t valid_chunks = 0) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(my_stage);
  const uint64_t n_tiles = (n_blocks + kTileBlocks - 1) / kTileBlocks;
  const uint64_t total_warps = (uint64_t)gridDim.x * warps_per_cta;
  const uint64_t total_chunks = valid_chunks ? valid_chunks : n_blocks * 8;
  // chunk (b, j) of a tile lands in slot b*8 + (j ^ (b & 7)); lane l copies chunks l, l+32, ...: b = 4i + (l >> 3)
  const uint32_t dst_even = stage_addr + (uint32_t)((lane >> 3) * 8 + ((lane & 7) ^ (lane >> 3))) * 16u;
  const uint32_t dst_odd = stage_addr + (uint32_t)((lane >> 3) * 8 + ((lane & 7) ^ (4 + (lane >> 3)))) * 16u;
  auto issue = [&](uint64_t t, int stage) {
    const uint4* src = lines + t * (kTileBlocks * 8) + lane;
    const uint32_t soff = (uint32_t)stage * (uint32_t)kTileBytes;
    if ((t + 1) * (kTileBlocks * 8) <= total_chunks) {
#pragma unroll
      for (int i = 0; i < 8; i++) cp_async16(((i & 1) ? dst_odd : dst_even) + soff + (uint32_t)i * 512u, src + 32 * i, 16u);
    } else {  // last, partial tile: missing blocks are zero-filled
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const bool ok = t * (kTileBlocks * 8) + (uint64_t)(32 * i + lane) < total_chunks;
        cp_async16(((i & 1) ? dst_odd : dst_even) + soff + (uint32_t)i * 512u, ok ? (src + 32 * i) : lines, ok ? 16u : 0u);
      }
    }
    cp_async_commit();
  };
  uint64_t t = (uint64_t)blockIdx.x * warps_per_cta + warp;
  int stage = 0;
  if (t < n_tiles) issue(t, 0);
  for (; t < n_tiles; t += total_warps) {
    const uint64_t next = t + total_warps;
    if (STAGES == 2 && next < n_tiles) {
      issue(next, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    uint32_t x[32];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint4 q = my_stage[stage * 256 + lane * 8 + (j ^ (lane & 7))];
      x[4 * j] = q.x; x[4 * j + 1] = q.y; x[4 * j + 2] = q.z; x[4 * j + 3] = q.w;
    }
    __syncwarp();  // the stage is overwritten by the prefetch issued in the next iteration
    if (STAGES == 1 && next < n_tiles) issue(next, 0);
    const uint64_t blk = t * kTileBlocks + lane;
    body(x, blk, blk < n_blocks);
    if (STAGES == 2) stage ^= 1;
  }
}

}  // namespace tile
}  // namespace mpc
