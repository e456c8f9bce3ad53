// adam.cu - one fused Adam step over the FLAT parameter vector (the optimizer step either side of the
// rollout). Replaces optimizer.step() of trainer.py:177 for torch.optim.Adam (amsgrad off): per parameter tensor PyTorch
// launches a handful of elementwise kernels (or its multi-tensor fused variants); the rollout's gradient already IS one
// flat vector (hdpo_rollout_bwd), so one launch updates parameters and both moment vectors. Same arithmetic and the same
// order of operations as torch/optim/adam.py::_single_tensor_adam (lerp for the first moment, addcmul for the second,
// sqrt(v) / sqrt(1 - beta2^t) + eps, step size lr / (1 - beta1^t)); the bias corrections are host doubles.
#include <cmath>

#include "hdpo_internal.cuh"

namespace hdpo {

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n, float w1,
                                                        float beta2, float w2, float eps, float weight_decay,
                                                        float step_size, float bc2_sqrt) {
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  const float pi = p[i];
  if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);  // grad.add(param, alpha=weight_decay)
  float mi = m[i];
  mi = mi + w1 * (gi - mi);  // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = v[i] * beta2 + w2 * gi * gi;  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  m[i] = mi;
  v[i] = vi;
  p[i] = pi - step_size * (mi / denom);  // param.addcdiv_(exp_avg, denom, value=-step_size)
}

}  // namespace hdpo

using namespace hdpo;

extern "C" int hdpo_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                              double beta1, double beta2, double eps, double weight_decay, int64_t step, void* stream) {
  HDPO_REQUIRE(params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad arguments");
  if (n == 0) return HDPO_OK;
  // hyper-parameters are doubles like PyTorch's Python floats: 1 - beta2 formed in float would be off by 1e-5 relative
  const double bc1 = 1.0 - std::pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - std::pow(beta2, static_cast<double>(step));
  const float step_size = static_cast<float>(lr / bc1);
  const float bc2_sqrt = static_cast<float>(std::sqrt(bc2));
  auto k = adam_step_kernel;
  HDPO_LAUNCH_PDL(k, static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, stream, params, grads, exp_avg, exp_avg_sq, n,
                  static_cast<float>(1.0 - beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta2),
                  static_cast<float>(eps), static_cast<float>(weight_decay), step_size, bc2_sqrt);
  HDPO_LAUNCH_OK();
  return HDPO_OK;
}
