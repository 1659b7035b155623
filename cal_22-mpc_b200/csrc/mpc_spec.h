// Registry of config-specialised kernels (thread-per-block mapping, tables folded at compile time).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mpc_capi.h"

namespace mpc {

struct SpecKernel {
  const char* name;  // config name the kernel was instantiated for (e.g. "P6")
  // true when `cfg` is exactly the config the kernel was compiled for
  bool (*matches)(const mpc_config_pod& cfg);
  cudaError_t (*launch)(const mpc_config_pod& cfg, const uint8_t* d_lines, uint64_t n_blocks, uint16_t* d_packed,
                        uint64_t* d_stats, const uint8_t* d_row_lut, uint32_t* d_sched, int sm_count, cudaStream_t stream);
  int lut_xor;  // which row-cost table the kernel expects (0 plain, 1 consecutive-XOR folded in, 2 first-plane-XOR folded in)
};

// table of the common encoder's cost of a non-zero 16-bit scan row (0 for the zero row), FPCModule.cpp:47-66, in the skewed
// layout of mpc_layout.h (kRowLutBytes bytes)
void build_row_cost_lut(uint8_t* lut, int lut_xor);

// true when the two configs describe the same computation (fields the kernels depend on)
bool spec_pod_equal(const mpc_config_pod& a, const mpc_config_pod& b);

// defined in mpc_spec_list.cu
extern const SpecKernel* const kSpecKernels[];
extern const int kNumSpecKernels;

// nullptr when no specialisation is linked in for this config
const SpecKernel* find_spec_kernel(const mpc_config_pod& cfg);

}  // namespace mpc
