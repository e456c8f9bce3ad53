// adjoint-kernel instantiations for input tile width KQ0 = 1 (MLP inputs padded to 4); see rollout_small_kernels.cuh
#include "rollout_small_kernels.cuh"

namespace hdpo {
namespace small {
HDPO_SMALL_BWD_INSTANCE(HDPO_ARCH_VANILLA_ONE_STORE, 1)
HDPO_SMALL_BWD_INSTANCE(HDPO_ARCH_VANILLA_SERIAL, 1)
}  // namespace small
}  // namespace hdpo
