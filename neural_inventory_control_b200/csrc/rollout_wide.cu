// rollout_wide.cu - placeholder until the tile-GEMM pipeline lands (next milestone).
#include "rollout_wide.cuh"

namespace hdpo {
namespace wide {

bool supported(const HdpoRolloutDesc*) { return false; }
size_t workspace_bytes(const HdpoRolloutDesc*) { return 0; }
int forward(const HdpoRolloutDesc*, const float*, const float*, const HdpoStatics*, const HdpoState*, float*, float*,
            float*, double*, HdpoState*, void*, size_t, void*) {
  set_error("wide rollout not built");
  return HDPO_E_INVALID;
}
int backward(const HdpoRolloutDesc*, const float*, const float*, const HdpoStatics*, float, float, float*, void*, size_t,
             void*) {
  set_error("wide rollout not built");
  return HDPO_E_INVALID;
}

}  // namespace wide
}  // namespace hdpo
