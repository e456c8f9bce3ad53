// philox.cu - K4 placeholder (filled in by the sampler milestone).
#include "hdpo_internal.cuh"

extern "C" int hdpo_philox_normal(float*, int32_t, int32_t, int32_t, int32_t, const float*, const float*, float, int32_t,
                                  uint64_t, uint64_t, void*) {
  hdpo::set_error("philox sampler not built");
  return HDPO_E_INVALID;
}
extern "C" int hdpo_philox_poisson(float*, int32_t, int32_t, int32_t, int32_t, const float*, uint64_t, uint64_t, void*) {
  hdpo::set_error("philox sampler not built");
  return HDPO_E_INVALID;
}
