// adjoint-kernel instantiations for input tile width KQ0 = 5 (MLP inputs padded to 20); see rollout_small_kernels.cuh
#include "rollout_small_kernels.cuh"

namespace hdpo {
namespace small {
HDPO_SMALL_BWD_INSTANCE(HDPO_ARCH_VANILLA_ONE_STORE, 5)
HDPO_SMALL_BWD_INSTANCE(HDPO_ARCH_VANILLA_SERIAL, 5)
}  // namespace small
}  // namespace hdpo
