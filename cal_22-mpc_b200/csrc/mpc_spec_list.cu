// Specialised kernels linked into the library (spec/spec_list.inc is written by tools/specgen).
#include "mpc_spec.h"

namespace mpc {
#define MPC_SPEC(n) extern const SpecKernel kSpec_##n;
#include "spec/spec_list.inc"
#undef MPC_SPEC
const SpecKernel* const kSpecKernels[] = {
#define MPC_SPEC(n) &kSpec_##n,
#include "spec/spec_list.inc"
#undef MPC_SPEC
    nullptr};
const int kNumSpecKernels = (int)(sizeof(kSpecKernels) / sizeof(kSpecKernels[0])) - 1;
}  // namespace mpc
