// Layout of the device statistics vector (uint64 words); shared by host code, nvcc and NVRTC translation units.
//   [0, K)            sum over stage-3 lines of sum_i |r_i|      per cluster k = selected + 1
//   [K, 2K)           sum over stage-3 lines of sum_i r_i^2
//   [2K, 2K + K*HB)   histogram  hist[k][size]        (K = MPC_MAX_MODULES + 1 = 17, HB = MPC_HIST_BINS = 1056)
// Counts and compressed-size totals are derived from the histogram on the host (mpc_stats_expand).
#pragma once
namespace mpc {
constexpr int kK = 16 + 1;
constexpr int kHB = 8 * 128 + 32;
constexpr unsigned long long kStatsWords = 2ull * kK + (unsigned long long)kK * kHB;
constexpr unsigned long long kResAbsOff = 0;
constexpr unsigned long long kResSqOff = kK;
constexpr unsigned long long kHistOff = 2 * kK;
// Row-cost table of the specialised kernels (column-major scans): entry of the 16-bit scan row (A << 8) | B sits at byte
// 260 * A + B.  The skew of 4 bytes per value of A spreads rows whose low byte is zero (common: the more significant byte of a
// small delta) over the shared-memory banks -- with the plain index A * 256 + B they all fall into bank 0 -- and costs nothing:
// the kernels form both indices of a word of two rows with one IDP.2A each (dot products with (260, 1)).
constexpr unsigned kRowLutSkew = 260;
constexpr unsigned kRowLutBytes = 66560;  // >= 260 * 255 + 255 + 1, a multiple of 16
// Per-warp deferral buffer of the plane-major kernels (mpc_spec.cuh): 32 blocks of 128 bytes + their indices + their winners + a counter.
constexpr unsigned kDeferBytesPerWarp = 32 * 128 + 32 * 4 + 32 * 4 + 128;
// A CUtensorMap (128 bytes, 64-byte aligned) as the kernels see it: opaque, so that the NVRTC build needs no cuda.h.
// Describes the dump as a [n_blocks][128 B] uint8 tensor with a 32 x 128 B box and the 128-byte swizzle
// (make_tile_tmap, mpc_jit.cpp).
struct alignas(64) TileTmap { unsigned long long opaque[16]; };
}  // namespace mpc
