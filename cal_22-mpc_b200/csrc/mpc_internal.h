// Internal declarations shared by the translation units of libmpc_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "mpc_capi.h"
#include "mpc_layout.h"

namespace mpc {

static_assert(kK == MPC_MAX_MODULES + 1 && kHB == MPC_HIST_BINS && kStatsWords == MPC_STATS_WORDS, "mpc_layout.h out of sync with mpc_capi.h");

// ---- generic kernel tables (one per PredComp module), built on the host from mpc_module_pod ---------
struct GenericModule {
  uint8_t xsrc[MPC_MAX_LINE];  // residue position j takes line byte xsrc[j]        (ResidueModule.cpp:28-39)
  uint8_t psrc[MPC_MAX_LINE];  // ... minus a prediction derived from line byte psrc[j]
  uint8_t pval[MPC_MAX_LINE];  // DiffTable byte (op 1) or shift distance as int8 (op 2)
  uint16_t sbit[8 * MPC_MAX_LINE];  // scan position i reads bit sbit[i] of the 8L-bit g array; 0xffff = none
  int32_t op;         // 0 = prediction is the source byte, 1 = + diff, 2 = shifted
  int32_t cxor;       // XORModule.consecutiveXOR
  int32_t root;       // RootIndex
  int32_t xidentity;  // xsrc[j] == j for all j >= 1 (root == 0)
  int32_t root_pred;  // line byte the predictor puts at position root (MAE/MSE only)
  int32_t pad[3];
};

struct GenericParams {
  int32_t line_size, words, rows;  // L, L/4, L/2
  int32_t num_modules, num_predcomp, first_predcomp, has_wordsame;
  int32_t hist_bins;               // 8L + 32
  int32_t enc_bits[MPC_MAX_MODULES + 1];
};

void build_generic_tables(const mpc_config_pod& cfg, GenericParams* params, GenericModule* mods);

cudaError_t launch_generic(const GenericParams& params, const GenericModule* d_mods, const uint8_t* d_lines,
                           uint64_t n_blocks, uint16_t* d_packed, uint64_t* d_stats, int sm_count,
                           cudaStream_t stream);

cudaError_t launch_synth(uint8_t* d_lines, uint64_t first_block, uint64_t n_blocks, uint64_t total_blocks, int kind,
                         uint64_t seed, cudaStream_t stream);

}  // namespace mpc
