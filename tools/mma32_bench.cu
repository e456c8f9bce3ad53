// tools/mma32_bench.cu - GPU-box microbenchmark: one 32-wide policy layer for the 32 rows of a warp,
//   out[r][n] = bias[n] + sum_k in[r][k] * W[n][k]      (r = lane's row in the SIMT form)
// (a) SIMT form of the small-net kernels (broadcast LDS.128 of Wt[k][n] + packed FFMA2, lane = row), against
// (b) warp-level tensor-core form: mma.sync.m16n8k8 tf32 with the 3xTF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulate),
//     A fragments from the same shared-memory rows, B fragments from a pre-split W[n][k] copy (stride 36).
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural_inventory_control_b200/csrc \
//               tools/mma32_bench.cu -o tools/_build/mma32_bench && tools/_build/mma32_bench
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

constexpr int H = 32, HS = 36;

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ float f4c(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// (a) lane = row
__device__ __forceinline__ void layer_simt(const float* __restrict__ Wt, const float* __restrict__ bias,
                                           const float* __restrict__ in, float* __restrict__ out) {
  float2 acc[H / 2];
#pragma unroll
  for (int n4 = 0; n4 < H / 4; ++n4) {
    const float4 bv = reinterpret_cast<const float4*>(bias)[n4];
    acc[2 * n4] = make_float2(bv.x, bv.y);
    acc[2 * n4 + 1] = make_float2(bv.z, bv.w);
  }
  for (int k4 = 0; k4 < H / 4; ++k4) {
    const float4 xv = reinterpret_cast<const float4*>(in)[k4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4* w = reinterpret_cast<const float4*>(Wt + (4 * k4 + kk) * H);
      const float xk = f4c(xv, kk);
      const float2 xx = make_float2(xk, xk);
#pragma unroll
      for (int n4 = 0; n4 < H / 4; ++n4) {
        const float4 wv = w[n4];
        ffma2(acc[2 * n4], make_float2(wv.x, wv.y), xx);
        ffma2(acc[2 * n4 + 1], make_float2(wv.z, wv.w), xx);
      }
    }
  }
#pragma unroll
  for (int n4 = 0; n4 < H / 4; ++n4)
    reinterpret_cast<float4*>(out)[n4] = make_float4(acc[2 * n4].x, acc[2 * n4].y, acc[2 * n4 + 1].x, acc[2 * n4 + 1].y);
}

#include "mma32.cuh"  // the shipped tensor-core forms (neural_inventory_control_b200/csrc)
__device__ __forceinline__ unsigned tf32_bits(float x) { return hdpo::mma32::tf32_bits(x); }
// (b) warp-cooperative: rows = 32 shared-memory rows (stride HS) of this warp; Whi / Wlo = W[n][k] with stride HS
__device__ __forceinline__ void layer_mma(const float* __restrict__ Whi, const float* __restrict__ Wlo,
                                          const float* __restrict__ bias, const float* __restrict__ in_rows,
                                          float* __restrict__ out_rows, int lane) {
  hdpo::mma32::layer<2>(Whi, Wlo, HS, bias, in_rows, HS, 4, out_rows, HS, lane);
}

template <int MODE>
__global__ void __launch_bounds__(256) bench_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                    const float* __restrict__ X, float* __restrict__ Y, int reps) {
  extern __shared__ __align__(16) float smem[];
  float* Wt = smem;                 // [k][n]
  float* Whi = Wt + H * H;          // [n][k] stride HS
  float* Wlo = Whi + H * HS;
  float* bs = Wlo + H * HS;
  float* rows = bs + H;             // per warp: in [32][HS], out [32][HS]
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
    const int n = i / H, k = i % H;
    const float w = W[i];
    Wt[k * H + n] = w;
    const float hi = __uint_as_float(tf32_bits(w));
    Whi[n * HS + k] = hi;
    Wlo[n * HS + k] = __uint_as_float(tf32_bits(w - hi));
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) bs[i] = bias[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* in = rows + warp * 2 * 32 * HS;
  float* out = in + 32 * HS;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  for (int k = 0; k < H; ++k) in[lane * HS + k] = X[(gw * 32 + lane) * H + k];
  __syncwarp();
  for (int r = 0; r < reps; ++r) {
    if (MODE == 0) layer_simt(Wt, bs, in + lane * HS, out + lane * HS);
    else layer_mma(Whi, Wlo, bs, in, out, lane);
    __syncwarp();
    float* tmp = in;  // ping-pong so that every repetition depends on the previous one (as consecutive layers do)
    in = out;
    out = tmp;
  }
  for (int k = 0; k < H; ++k) Y[(gw * 32 + lane) * H + k] = in[lane * HS + k];
}

int main() {
  const int ctas = 148 * 2, warps = 8, rows = ctas * warps * 32;
  std::vector<float> W(H * H), b(H), X(rows * H);
  srand(1);
  for (auto& v : W) v = (rand() / (float)RAND_MAX * 2 - 1) * 0.5f / sqrtf(H);  // contraction: repeated layers stay bounded
  for (auto& v : b) v = rand() / (float)RAND_MAX - 0.5f;
  for (auto& v : X) v = rand() / (float)RAND_MAX * 2 - 1;
  float *dW, *db, *dX, *dY;
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&db, b.size() * 4);
  cudaMalloc(&dX, X.size() * 4);
  cudaMalloc(&dY, X.size() * 4);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (H * H + 2 * H * HS + H + warps * 2 * 32 * HS) * sizeof(float);
  cudaFuncSetAttribute(bench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // numerics: one layer against a double reference
  std::vector<float> Y(X.size());
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 0) bench_kernel<0><<<ctas, warps * 32, smem>>>(dW, db, dX, dY, 1);
    else bench_kernel<1><<<ctas, warps * 32, smem>>>(dW, db, dX, dY, 1);
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, mag = 0;
    for (int r = 0; r < 4096; ++r)
      for (int n = 0; n < H; ++n) {
        double s = b[n];
        for (int k = 0; k < H; ++k) s += (double)X[r * H + k] * W[n * H + k];
        err = fmax(err, fabs(s - Y[r * H + n]));
        mag = fmax(mag, fabs(s));
      }
    printf("mode %d (%s): max abs err %.3e (max |y| %.3f), launch status %s\n", mode, mode ? "mma.sync 3xTF32" : "SIMT FFMA2",
           err, mag, cudaGetErrorString(cudaGetLastError()));
  }
  const int reps = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) {
      cudaEventRecord(e0);
      if (mode == 0) bench_kernel<0><<<ctas, warps * 32, smem>>>(dW, db, dX, dY, reps);
      else bench_kernel<1><<<ctas, warps * 32, smem>>>(dW, db, dX, dY, reps);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double layers = (double)ctas * warps * reps;
    printf("mode %d: %.3f ms, %.1f ns per dependent layer of a warp (16 warps per SM), %.2f TFLOP/s algorithmic\n", mode, ms,
           ms * 1e6 / reps, 2.0 * 32 * 32 * 32 * layers / (ms * 1e-3) / 1e12);
  }
  return 0;
}
