// Config compiler interface (mpc_specgen.cpp): config -> source of a specialised kernel schedule.
#pragma once
#include <cstddef>
#include <string>

#include "mpc_capi.h"

namespace mpc {

struct SpecTraits {
  bool eligible = false;
  int line_size = 128;           // bytes per line (32 / 64 / 128); a tile of 4 KiB holds 32 * (128 / line_size) lines
  bool use_lut = false;          // column-major modules: 64 KiB shared-memory row-cost table
  int lut_xor = 0;               // 0 plain table, 1 / 2: consecutive / first-plane XOR stage folded into the table
  int warps = 8;                 // warps per CTA
  int stages = 2;                // shared-memory tile stages per warp (1: the registers are the second buffer)
  int min_ctas = 2;              // __launch_bounds__ second argument
  bool skip_zero_groups = true;  // encoder branches around groups of eight zero rows
  bool tma = false;              // tiles arrive by one TMA tensor copy per warp (cp.async.bulk.tensor + mbarrier) instead of 8 cp.async per lane
  int queue_cap = 0;             // entries per regrouping queue (mpc_spec.cuh), 0 = none
  bool fused_encode = false;     // every module's winner pass runs into its own copy of the row classifier
  bool adaptive_encode = false;  // fused passes for warps whose lanes agree on the module, the shared classifier otherwise
  bool defer = false;            // pm2: per-warp deferral buffer (blocks whose winner is not the last module are finished in full batches)
  bool pm2 = false;              // all modules plane-major with one scan family: one residue pass per module, shared scoring / statistics / classifier code
  size_t smem_bytes = 0;         // dynamic shared memory of one CTA
};

// true when a specialised kernel can be generated for cfg (lineSize 32 / 64 / 128, every scan column- or plane-major)
bool spec_eligible(const mpc_config_pod& cfg, std::string* why);
SpecTraits spec_traits(const mpc_config_pod& cfg);
// jit = false: translation unit for the ahead-of-time build (registers itself as kSpec_<name>);
// jit = true: NVRTC translation unit exposing `extern "C" __global__ mpc_jit_kernel`.  Empty string + *why when not eligible.
std::string generate_spec_source(const mpc_config_pod& cfg, const std::string& name, bool jit, std::string* why);

// test hook: the residue statements of every PredComp module as host C++ (tests/test_specgen_residues.py); empty + *why when not eligible
std::string generate_residue_probe(const mpc_config_pod& cfg, std::string* why);

}  // namespace mpc
