// Run-time instantiation of the specialised kernel (NVRTC) for configs that were not compiled into the library.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "mpc_capi.h"
#include "mpc_layout.h"
#include "mpc_specgen.h"

namespace mpc {

struct JitKernel;  // loaded module + function + launch traits

// Source generation + NVRTC compilation only (no device needed).  Returns MPC_OK and the cubin, or an error + log.
int jit_compile(const mpc_config_pod& cfg, std::vector<char>* cubin, SpecTraits* traits, std::string* log);
// Compile (or fetch from the MPC_JIT_CACHE_DIR disk cache) and load into the current device's primary context.
JitKernel* jit_create(const mpc_config_pod& cfg, std::string* log);
void jit_destroy(JitKernel* k);
int jit_lut_xor(const JitKernel* k);
cudaError_t jit_launch(JitKernel* k, const uint8_t* d_lines, uint64_t n_blocks, uint16_t* d_packed, uint64_t* d_stats,
                       const uint8_t* d_row_lut, uint32_t* d_sched, int sm_count, cudaStream_t stream);

// Tensor map of a dump for the TMA tile loader of the specialised kernels: uint8 [n_blocks][128], box 32 x 128 B,
// SWIZZLE_128B (the layout the per-thread 16-byte reads are conflict-free on), out-of-range rows read as zero.
// cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint like the other driver entry points.
cudaError_t make_tile_tmap(TileTmap* out, const uint8_t* d_lines, uint64_t n_blocks);

}  // namespace mpc
