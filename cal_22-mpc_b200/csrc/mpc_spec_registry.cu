// Specialised-kernel registry: finds the compiled-in kernel whose config equals the context's config.
#include <cstring>

#include "mpc_spec.h"
#include "mpc_device.cuh"
#include "mpc_layout.h"

namespace mpc {

bool spec_pod_equal(const mpc_config_pod& a, const mpc_config_pod& b) {
  if (a.line_size != b.line_size || a.num_modules != b.num_modules || a.has_wordsame != b.has_wordsame ||
      a.first_predcomp != b.first_predcomp)
    return false;
  const int n = a.num_modules, L = a.line_size;
  if (n < 1 || n > MPC_MAX_MODULES || L < 1 || L > MPC_MAX_LINE) return false;
  for (int i = 0; i <= n; i++)
    if (a.enc_bits[i] != b.enc_bits[i]) return false;
  for (int i = 0; i < n; i++) {
    const mpc_module_pod &x = a.modules[i], &y = b.modules[i];
    if (x.kind != y.kind) return false;
    if (x.kind != MPC_MOD_PREDCOMP) continue;
    if (x.predictor != y.predictor || x.root != y.root || x.consecutive_xor != y.consecutive_xor ||
        x.table_size != y.table_size)
      return false;
    if (x.table_size < 0 || x.table_size > 8 * L) return false;
    if (memcmp(x.scan_row, y.scan_row, (size_t)x.table_size) || memcmp(x.scan_col, y.scan_col, (size_t)x.table_size))
      return false;
    if (x.predictor == MPC_PRED_DIFF || x.predictor == MPC_PRED_WEIGHT) {
      for (int j = 0; j < L; j++) {
        if (j == x.root) continue;
        if (x.base[j] != y.base[j]) return false;
        if (x.predictor == MPC_PRED_DIFF && x.diff[j] != y.diff[j]) return false;
        if (x.predictor == MPC_PRED_WEIGHT && x.shift[j] != y.shift[j]) return false;
      }
    }
  }
  return true;
}

void build_row_cost_lut(uint8_t* lut, int lut_xor) {
  memset(lut, 0, kRowLutBytes);
  for (uint32_t v = 0; v < 65536; v++) {
    uint32_t row = v;  // two residue bytes; with lut_xor the XOR stage (XORModule.cpp:9-20) is applied here, per byte
    if (lut_xor == 1) row = mpcdev::xor_planes_consecutive(v, 0) & 0xffffu;
    else if (lut_xor == 2) row = mpcdev::xor_planes_first(v, 0) & 0xffffu;
    uint32_t nz;
    lut[kRowLutSkew * (v >> 8) + (v & 0xffu)] = (uint8_t)(mpcdev::row2_cost(row, &nz) & 0xffu);  // skewed layout, mpc_layout.h
  }
}

const SpecKernel* find_spec_kernel(const mpc_config_pod& cfg) {
  for (int i = 0; i < kNumSpecKernels; i++)
    if (kSpecKernels[i]->matches(cfg)) return kSpecKernels[i];
  return nullptr;
}

}  // namespace mpc
