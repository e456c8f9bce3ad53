// adjoint-kernel instantiations for input tile width KQ0 = 2 (MLP inputs padded to 8); see rollout_small_kernels.cuh
#include "rollout_small_kernels.cuh"

namespace hdpo {
namespace small {
HDPO_SMALL_BWD_INSTANCE(HDPO_ARCH_VANILLA_ONE_STORE, 2)
HDPO_SMALL_BWD_INSTANCE(HDPO_ARCH_VANILLA_SERIAL, 2)
}  // namespace small
}  // namespace hdpo
