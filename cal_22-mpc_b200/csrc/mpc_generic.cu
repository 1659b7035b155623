// Generic MPC kernel: one warp per block, every table read from shared memory at run time.
//
// This is the always-available path: it accepts every config mpc_config_validate() accepts
// (arbitrary predictor base tables, arbitrary 8L-entry scan permutations, L = 32/64/128) and is the
// mapping BASELINE.json's north_star prescribes -- lane l holds word l of the block, zero / repeat
// detection and the row tests are warp votes, sums are warp reductions.  Specialised kernels
// (mpc_spec.cuh) overtake it on throughput; it stays as the reference point for that comparison
// and as the fallback for configs without a specialisation.
//
// Reference path restated here (src/compressor/): VPC.cpp:22-70 (cascade), 332-415 (checks,
// selection, raw fallback), 417-443 (residue stats); VPCmodules/* (predict, residue, bit planes,
// xor, scan); FPCModule.cpp:19-158 (common encoder).
#include <cuda_runtime.h>
#include <stdint.h>

#include "mpc_device.cuh"
#include "mpc_internal.h"

namespace mpc {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;

struct WarpScratch {
  uint32_t x[MPC_MAX_LINE / 4];  // the block, byte-addressable for predictor gathers
  uint32_t g[MPC_MAX_LINE / 4];  // xor-ed residue bytes, bit-addressable for the scan gather
};

__device__ __forceinline__ uint32_t ld_byte(const uint32_t* words, uint32_t idx) {
  return reinterpret_cast<const uint8_t*>(words)[idx];
}

// ScanModule::ProcessLine for one output row (ScanModule.cpp:14-20): 16 table-driven bit reads.
// Scan position k of the row lands in bit 15-k.
__device__ __forceinline__ uint32_t gather_row(const uint16_t* sbit_row, const uint32_t* g) {
  uint32_t v = 0;
  const uint32_t* pairs = reinterpret_cast<const uint32_t*>(sbit_row);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t pr = pairs[k];
    uint32_t i0 = pr & 0xffffu, i1 = pr >> 16;
    uint32_t b0 = (i0 == 0xffffu) ? 0u : ((g[i0 >> 5] >> (i0 & 31u)) & 1u);
    uint32_t b1 = (i1 == 0xffffu) ? 0u : ((g[i1 >> 5] >> (i1 & 31u)) & 1u);
    v |= b0 << (15 - 2 * k);
    v |= b1 << (14 - 2 * k);
  }
  return v;
}

__global__ void __launch_bounds__(kThreads)
mpc_generic_kernel(const GenericParams P, const GenericModule* __restrict__ g_mods,
                   const uint32_t* __restrict__ lines, uint64_t n_blocks, uint16_t* __restrict__ packed,
                   unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [GenericModule x num_predcomp][hist u32 x K*HB][res u64 x 2K][WarpScratch x warps]
  const int K = P.num_modules + 1;
  const int HB = P.hist_bins;
  GenericModule* s_mods = reinterpret_cast<GenericModule*>(smem_raw);
  unsigned long long* s_res = reinterpret_cast<unsigned long long*>(s_mods + P.num_predcomp);
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_res + 2 * K);
  WarpScratch* s_scr = reinterpret_cast<WarpScratch*>(s_hist + K * HB);

  {  // stage tables, clear CTA statistics
    const uint32_t* src = reinterpret_cast<const uint32_t*>(g_mods);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_mods);
    const int nw = (int)(sizeof(GenericModule) / 4) * P.num_predcomp;
    for (int i = threadIdx.x; i < nw; i += kThreads) dst[i] = src[i];
    for (int i = threadIdx.x; i < K * HB; i += kThreads) s_hist[i] = 0;
    for (int i = threadIdx.x; i < 2 * K; i += kThreads) s_res[i] = 0;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  WarpScratch& scr = s_scr[warp];
  const int W = P.words;          // words (= active lanes) per block
  const int R = P.rows;           // scan rows per block
  const int R2 = R >> 1;          // rows handled per halfword slot
  const uint32_t full = 0xffffffffu;
  const uint32_t r2mask = (R2 >= 32) ? full : ((1u << R2) - 1u);
  const bool active = lane < W;
  const uint32_t keep = (lane == 0) ? 0xffu : 0u;

  const uint64_t total_warps = (uint64_t)gridDim.x * kWarpsPerCta;
  uint64_t blk = (uint64_t)blockIdx.x * kWarpsPerCta + warp;
  uint32_t x_next = 0;
  if (blk < n_blocks && active) x_next = lines[blk * W + lane];

  for (; blk < n_blocks; blk += total_warps) {
    const uint32_t x = x_next;
    {  // prefetch the next block of this warp
      uint64_t nb = blk + total_warps;
      if (nb < n_blocks && active) x_next = lines[nb * W + lane];
    }
    int sel;
    uint32_t size;
    bool stage3 = false;
    uint32_t sum_abs = 0, sum_sq = 0;

    const bool all_zero = __all_sync(full, x == 0u);  // AllZeroModule.cpp:7-15
    const uint32_t x0 = __shfl_sync(full, x, 0);
    if (all_zero) {
      sel = 0;
      size = (uint32_t)P.enc_bits[1];  // VPC.cpp:343
    } else if (P.has_wordsame && __all_sync(full, !active || x == x0)) {  // AllWordSameModule.cpp:7-21
      sel = 1;
      size = 32u + (uint32_t)P.enc_bits[2];  // VPC.cpp:358
    } else {
      // ---- VPC::checkOtherPatterns, VPC.cpp:366-415 ----
      stage3 = true;
      if (active) scr.x[lane] = x;
      __syncwarp();
      int best = -1;
      uint32_t bestz = 0, best_rows = 0, best_r = 0, best_extra = 0;
      uint64_t best_zmask = 0;
      for (int m = 0; m < P.num_predcomp; m++) {
        const GenericModule& gm = s_mods[m];
        uint32_t r = 0;
        if (active) {
          uint32_t xg = x;
          if (!gm.xidentity) {
            uint32_t xs = reinterpret_cast<const uint32_t*>(gm.xsrc)[lane];
            xg = ld_byte(scr.x, xs & 0xff) | (ld_byte(scr.x, (xs >> 8) & 0xff) << 8) |
                 (ld_byte(scr.x, (xs >> 16) & 0xff) << 16) | (ld_byte(scr.x, xs >> 24) << 24);
          }
          const uint32_t ps = reinterpret_cast<const uint32_t*>(gm.psrc)[lane];
          uint32_t pg = ld_byte(scr.x, ps & 0xff) | (ld_byte(scr.x, (ps >> 8) & 0xff) << 8) |
                        (ld_byte(scr.x, (ps >> 16) & 0xff) << 16) | (ld_byte(scr.x, ps >> 24) << 24);
          if (gm.op == 1) {  // DiffBasePredictor, PredictorModule.cpp:104
            pg = mpcdev::add_u8x4(pg, reinterpret_cast<const uint32_t*>(gm.pval)[lane]);
          } else if (gm.op == 2) {  // WeightBasePredictor, PredictorModule.cpp:58-62
            const uint32_t pv = reinterpret_cast<const uint32_t*>(gm.pval)[lane];
            uint32_t out = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
              int s = (int)(int8_t)((pv >> (8 * q)) & 0xff);
              uint32_t b = (pg >> (8 * q)) & 0xffu;
              uint32_t v = (s < 0) ? (b >> (-s)) : ((b << s) & 0xffu);
              out |= v << (8 * q);
            }
            pg = out;
          }
          r = mpcdev::sub_u8x4(xg, pg);                                   // ResidueModule.cpp:34
          if (lane == 0) r = (r & ~0xffu) | ld_byte(scr.x, (uint32_t)gm.root);  // root first, ResidueModule.cpp:26-27
          const uint32_t g = gm.cxor ? mpcdev::xor_planes_consecutive(r, keep) : mpcdev::xor_planes_first(r, keep);
          scr.g[lane] = g;
        }
        __syncwarp();
        uint32_t v0 = 0, v1 = 0;
        if (lane < R2) {
          v0 = gather_row(gm.sbit + 16 * lane, scr.g);
          v1 = gather_row(gm.sbit + 16 * (lane + R2), scr.g);
        }
        const uint32_t zlo = __ballot_sync(full, v0 == 0u) & r2mask;
        const uint32_t zhi = __ballot_sync(full, v1 == 0u) & r2mask;
        const uint64_t zmask = (uint64_t)zlo | ((uint64_t)zhi << R2);
        const uint32_t z = mpcdev::leading_zero_rows(zmask, (uint32_t)R);  // VPC.cpp:378-387
        if (bestz <= z) {                                                  // ties -> later module, VPC.cpp:389
          best = P.first_predcomp + m;
          bestz = z;
          best_rows = v0 | (v1 << 16);
          best_zmask = zmask;
          best_r = r;
          // residue of the root position itself, which MAE/MSE include but the residue line does not
          best_extra = (ld_byte(scr.x, (uint32_t)gm.root) - ld_byte(scr.x, (uint32_t)gm.root_pred)) & 0xffu;
        }
        __syncwarp();
      }
      // ---- common encoder, FPCModule.cpp:19-85 ----
      uint32_t nz;
      uint32_t c2 = mpcdev::row2_cost(best_rows, &nz);
      uint32_t cost = __reduce_add_sync(full, (c2 & 0xffffu) + (c2 >> 16));
      if (best >= 0) cost += mpcdev::zero_run_cost(best_zmask);
      // ---- raw fallback + encoding bits, VPC.cpp:398-407 ----
      uint32_t rr;
      if (cost < 8u * (uint32_t)P.line_size) {
        sel = best;
        size = cost;
        rr = (lane == 0) ? ((best_r & ~0xffu) | best_extra) : best_r;
      } else {
        sel = -1;
        size = 8u * (uint32_t)P.line_size;
        rr = x;  // VPC.cpp:429-439
      }
      if (best < 0) rr = x;
      size += (uint32_t)P.enc_bits[sel + 1];
      sum_abs = __reduce_add_sync(full, mpcdev::sum_u8x4(rr));   // ResidueModule.cpp:43-57
      sum_sq = __reduce_add_sync(full, mpcdev::sumsq_u8x4(rr));  // ResidueModule.cpp:59-73
    }
    if (lane == 0) {
      const int k = sel + 1;
      atomicAdd(&s_hist[k * HB + (int)size], 1u);  // VPC.h:49-60
      if (stage3) {                                // VPC.h:62-76
        atomicAdd(&s_res[k], (unsigned long long)sum_abs);
        atomicAdd(&s_res[K + k], (unsigned long long)sum_sq);
      }
      if (packed) packed[blk] = (uint16_t)(size | ((uint32_t)k << 11));
    }
  }
  __syncthreads();
  // flush CTA statistics
  for (int i = threadIdx.x; i < K * HB; i += kThreads) {
    uint32_t c = s_hist[i];
    if (c) {
      int k = i / HB, s = i - k * HB;
      atomicAdd(&stats[kHistOff + (size_t)k * kHB + s], (unsigned long long)c);
    }
  }
  for (int i = threadIdx.x; i < K; i += kThreads) {
    if (s_res[i]) atomicAdd(&stats[kResAbsOff + i], s_res[i]);
    if (s_res[K + i]) atomicAdd(&stats[kResSqOff + i], s_res[K + i]);
  }
}

}  // namespace

void build_generic_tables(const mpc_config_pod& cfg, GenericParams* params, GenericModule* mods) {
  const int L = cfg.line_size;
  GenericParams& P = *params;
  P.line_size = L;
  P.words = L / 4;
  P.rows = L / 2;
  P.num_modules = cfg.num_modules;
  P.first_predcomp = cfg.first_predcomp;
  P.has_wordsame = cfg.has_wordsame;
  P.num_predcomp = cfg.num_modules - cfg.first_predcomp;
  P.hist_bins = 8 * L + 32;
  for (int i = 0; i <= MPC_MAX_MODULES; i++) P.enc_bits[i] = cfg.enc_bits[i];
  // byte-plane transposed copy used by ConsecutiveBasePredictor (PredictorModule.cpp:143-155)
  int tperm[MPC_MAX_LINE];
  {
    int idx = 0;
    for (int plane = 3; plane >= 0; plane--)
      for (int i = plane; i < L; i += 4) tperm[idx++] = i;
  }
  for (int m = 0; m < P.num_predcomp; m++) {
    const mpc_module_pod& src = cfg.modules[cfg.first_predcomp + m];
    GenericModule& gm = mods[m];
    gm = GenericModule();
    gm.root = src.root;
    gm.cxor = src.consecutive_xor;
    gm.xidentity = (src.root == 0);
    gm.root_pred = (src.predictor == MPC_PRED_CONSEC) ? tperm[src.root] : src.root;
    gm.op = 0;
    bool any_diff = false;
    for (int j = 0; j < L; j++) {
      const int i = (j == 0) ? src.root : (j <= src.root ? j - 1 : j);  // ResidueModule.cpp:28-39
      gm.xsrc[j] = (uint8_t)i;
      uint8_t ps = 0, pv = 0;
      switch (src.predictor) {
        case MPC_PRED_ONE: ps = (uint8_t)src.root; break;                         // PredictorModule.cpp:113-130
        case MPC_PRED_CONSEC: ps = (uint8_t)tperm[i > 0 ? i - 1 : 0]; break;       // PredictorModule.cpp:157-171
        case MPC_PRED_DIFF: ps = src.base[i]; pv = src.diff[i]; any_diff |= (pv != 0); break;
        case MPC_PRED_WEIGHT: ps = src.base[i]; pv = (uint8_t)src.shift[i]; break;
      }
      if (j == 0) { ps = (uint8_t)src.root; pv = 0; }  // overwritten by the root byte anyway
      gm.psrc[j] = ps;
      gm.pval[j] = pv;
    }
    if (src.predictor == MPC_PRED_DIFF && any_diff) gm.op = 1;
    if (src.predictor == MPC_PRED_WEIGHT) gm.op = 2;
    for (int i = 0; i < 8 * MPC_MAX_LINE; i++) gm.sbit[i] = 0xffffu;
    for (int i = 0; i < src.table_size; i++)
      gm.sbit[i] = (uint16_t)(src.scan_col[i] * 8 + (7 - src.scan_row[i]));  // plane b = bit 7-b, BitplaneModule.cpp:31
  }
}

cudaError_t launch_generic(const GenericParams& params, const GenericModule* d_mods, const uint8_t* d_lines,
                           uint64_t n_blocks, uint16_t* d_packed, uint64_t* d_stats, int sm_count,
                           cudaStream_t stream) {
  if (n_blocks == 0) return cudaSuccess;
  const int K = params.num_modules + 1;
  size_t smem = sizeof(GenericModule) * (size_t)params.num_predcomp + sizeof(unsigned long long) * 2 * K +
                sizeof(uint32_t) * (size_t)K * params.hist_bins + sizeof(WarpScratch) * kWarpsPerCta;
  // the opt-in is a per-device (per-context) attribute: set it on every launch, as launch_spec does
  cudaError_t e = cudaFuncSetAttribute(mpc_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mpc_generic_kernel, kThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  uint64_t want = (n_blocks + kWarpsPerCta - 1) / kWarpsPerCta;
  uint64_t grid = (uint64_t)sm_count * per_sm;  // persistent: a whole number of CTAs per SM
  if (grid > want) grid = want;
  mpc_generic_kernel<<<(unsigned)grid, kThreads, smem, stream>>>(
      params, d_mods, reinterpret_cast<const uint32_t*>(d_lines), n_blocks, d_packed,
      reinterpret_cast<unsigned long long*>(d_stats));
  return cudaGetLastError();
}

}  // namespace mpc
