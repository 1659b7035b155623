// Specialised-kernel registry.  Instantiations register themselves in kSpecKernels.
#include "mpc_spec.h"

namespace mpc {

const SpecKernel* find_spec_kernel(const mpc_config_pod& cfg) {
  for (int i = 0; i < kNumSpecKernels; i++)
    if (kSpecKernels[i]->matches(cfg)) return kSpecKernels[i];
  return nullptr;
}

}  // namespace mpc
