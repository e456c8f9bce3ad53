// cuda_emu.h - TEST INFRASTRUCTURE ONLY: a minimal CUDA-on-host-threads shim.
//
// Lets the SIMT kernel sources under neural_inventory_control_b200/csrc/ be compiled with g++ (-DHDPO_EMU) and
// executed in the CPU-only build container, so their arithmetic (forward recurrences, reverse-time adjoint,
// block-cooperative weight-gradient reductions, barriers) can be checked against the oracle before GPU time is
// spent. One std::thread per CUDA thread, blocks executed one after another, std::barrier for __syncthreads /
// warp shuffles. "Device" pointers are plain host pointers. Nothing here is performance relevant and the
// product package never loads the emulated library.
#pragma once

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu {
  unsigned x, y, z;
};
struct float2 {
  float x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };

namespace hdpo_emu {

struct BlockCtx {
  std::barrier<>* block_bar = nullptr;
  std::vector<std::unique_ptr<std::barrier<>>>* warp_bars = nullptr;
  unsigned char* dyn_smem = nullptr;
  uint64_t* xchg = nullptr;  // one slot per thread, for shuffles
  unsigned nthreads = 0;
};

inline thread_local uint3_emu t_threadIdx, t_blockIdx;
inline thread_local unsigned t_linear = 0;
inline thread_local BlockCtx* t_ctx = nullptr;
inline dim3 g_blockDim, g_gridDim;

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F f) {
  g_blockDim = block;
  g_gridDim = grid;
  const unsigned n = block.x * block.y * block.z;
  std::vector<unsigned char> dyn(smem + 64);
  unsigned char* dyn_aligned = dyn.data() + ((64 - (reinterpret_cast<uintptr_t>(dyn.data()) & 63)) & 63);
  std::vector<uint64_t> xchg(n);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(n);
        std::vector<std::unique_ptr<std::barrier<>>> wbars;
        for (unsigned w = 0; w * 32 < n; ++w) wbars.emplace_back(new std::barrier<>(std::min(32u, n - w * 32)));
        BlockCtx ctx{&bar, &wbars, dyn_aligned, xchg.data(), n};
        std::vector<std::thread> th;
        th.reserve(n);
        for (unsigned i = 0; i < n; ++i) {
          th.emplace_back([&, i]() {
            t_linear = i;
            t_threadIdx = {i % block.x, (i / block.x) % block.y, i / (block.x * block.y)};
            t_blockIdx = {bx, by, bz};
            t_ctx = &ctx;
            f();
            bar.arrive_and_drop();  // exited threads no longer take part in __syncthreads
          });
        }
        for (auto& t : th) t.join();
      }
}

inline void* dyn_smem() { return t_ctx->dyn_smem; }

template <class T>
inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  BlockCtx* c = t_ctx;
  const unsigned warp = t_linear / 32;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  c->xchg[t_linear] = bits;
  (*c->warp_bars)[warp]->arrive_and_wait();
  uint64_t got = c->xchg[warp * 32 + (static_cast<unsigned>(src_lane) & 31)];
  (*c->warp_bars)[warp]->arrive_and_wait();
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}

}  // namespace hdpo_emu

#define threadIdx (::hdpo_emu::t_threadIdx)
#define blockIdx (::hdpo_emu::t_blockIdx)
#define blockDim (::hdpo_emu::g_blockDim)
#define gridDim (::hdpo_emu::g_gridDim)

static inline void __syncthreads() { ::hdpo_emu::t_ctx->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  (*::hdpo_emu::t_ctx->warp_bars)[::hdpo_emu::t_linear / 32]->arrive_and_wait();
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return ::hdpo_emu::shfl_from(v, static_cast<int>(::hdpo_emu::t_linear & 31) ^ lane_mask);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int delta) {
  int lane = static_cast<int>(::hdpo_emu::t_linear & 31);
  int src = lane + delta;
  T r = ::hdpo_emu::shfl_from(v, src > 31 ? lane : src);
  return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
  return ::hdpo_emu::shfl_from(v, src);
}
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}

static inline float atomicAdd(float* addr, float v) {
  uint32_t* a = reinterpret_cast<uint32_t*>(addr);
  uint32_t old = __atomic_load_n(a, __ATOMIC_RELAXED), nw;
  float f;
  do {
    std::memcpy(&f, &old, 4);
    float s = f + v;
    std::memcpy(&nw, &s, 4);
  } while (!__atomic_compare_exchange_n(a, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return f;
}
static inline double atomicAdd(double* addr, double v) {
  uint64_t* a = reinterpret_cast<uint64_t*>(addr);
  uint64_t old = __atomic_load_n(a, __ATOMIC_RELAXED), nw;
  double f;
  do {
    std::memcpy(&f, &old, 8);
    double s = f + v;
    std::memcpy(&nw, &s, 8);
  } while (!__atomic_compare_exchange_n(a, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return f;
}
static inline int atomicAdd(int* addr, int v) { return __atomic_fetch_add(addr, v, __ATOMIC_RELAXED); }
static inline int atomicOr(int* addr, int v) { return __atomic_fetch_or(addr, v, __ATOMIC_RELAXED); }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned __float_as_uint(float f) {
  unsigned u;
  std::memcpy(&u, &f, 4);
  return u;
}
static inline float __uint_as_float(unsigned u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<uint64_t>(a) * b) >> 32); }
static inline float __fdividef(float a, float b) { return a / b; }

// runtime-API subset used by the host wrappers
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  std::memset(p, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memcpy(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

#define HDPO_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(::hdpo_emu::dyn_smem())
#define HDPO_LAUNCH(kfn, grid, block, smem, stream, ...)                                         \
  do {                                                                                           \
    ::hdpo_emu::launch(dim3(grid), dim3(block), (smem), [=]() { kfn(__VA_ARGS__); });            \
    ::hdpo::count_launch();                                                                      \
  } while (0)
