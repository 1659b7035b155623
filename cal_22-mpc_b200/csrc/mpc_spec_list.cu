// List of specialised kernels linked into the library (none yet: the generic kernel serves all configs).
#include "mpc_spec.h"

namespace mpc {
const SpecKernel* const kSpecKernels[] = {nullptr};
const int kNumSpecKernels = 0;
}  // namespace mpc
