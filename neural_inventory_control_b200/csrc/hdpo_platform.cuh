// hdpo_platform.cuh - thin platform layer for the SIMT kernels.
//
// Product build (nvcc, sm_100a): plain CUDA. Kernels are launched with HDPO_LAUNCH on the caller's stream.
// Test build (-DHDPO_EMU, g++, no GPU): tests/emu/cuda_emu.h maps the CUDA execution model onto host threads
// (one std::thread per CUDA thread, std::barrier for __syncthreads) so the *same kernel sources* can be checked
// against the oracle in the CPU-only build container. The emulated library is test infrastructure: it is built
// by tests/emu/build_emu.py into tests/emu/_build/ and is never loaded by the product package.
#pragma once

#include <cstddef>
#include <cstdint>

#ifdef HDPO_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace hdpo {

void count_launch();  // capi.cu

#ifndef HDPO_EMU
#define HDPO_DYN_SMEM(type, name)                                  \
  extern __shared__ __align__(16) unsigned char hdpo_dyn_smem_[];  \
  type* name = reinterpret_cast<type*>(hdpo_dyn_smem_)

// kfn must be a variable holding the kernel (so template commas are fine): auto kfn = kernel<A,B>;
#define HDPO_LAUNCH(kfn, grid, block, smem, stream, ...)                          \
  do {                                                                            \
    kfn<<<(grid), (block), (smem), reinterpret_cast<cudaStream_t>(stream)>>>(__VA_ARGS__); \
    ::hdpo::count_launch();                                                       \
  } while (0)
#endif

// Programmatic dependent launch (PDL): a kernel launched with HDPO_LAUNCH_PDL may be scheduled while its predecessor
// in the stream is still draining; it MUST call pdl_wait() before touching memory the predecessor wrote. This hides
// the launch latency of the long chains of short dependent kernels in the wide rollout (5 launches per period).
#ifndef HDPO_EMU
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define HDPO_LAUNCH_PDL(kfn, grid, block, smem, stream_, ...)                                 \
  do {                                                                                        \
    cudaLaunchConfig_t hdpo_cfg_{};                                                           \
    hdpo_cfg_.gridDim = dim3(grid);                                                           \
    hdpo_cfg_.blockDim = dim3(block);                                                         \
    hdpo_cfg_.dynamicSmemBytes = (smem);                                                      \
    hdpo_cfg_.stream = reinterpret_cast<cudaStream_t>(stream_);                               \
    cudaLaunchAttribute hdpo_attr_[1];                                                        \
    hdpo_attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                    \
    hdpo_attr_[0].val.programmaticStreamSerializationAllowed = 1;                             \
    hdpo_cfg_.attrs = hdpo_attr_;                                                             \
    hdpo_cfg_.numAttrs = 1;                                                                   \
    cudaLaunchKernelEx(&hdpo_cfg_, kfn, __VA_ARGS__);                                         \
    ::hdpo::count_launch();                                                                   \
  } while (0)
#else
inline void pdl_wait() {}
#define HDPO_LAUNCH_PDL(kfn, grid, block, smem, stream, ...) HDPO_LAUNCH(kfn, grid, block, smem, stream, __VA_ARGS__)
#endif

// Optional device-side trace (tools/trace_step.py): one record per CTA {start ns, end ns, smid | tag << 16, block id}
// appended to a caller-provided buffer (buf[0] = record counter, records from buf[4]). Null buffer = disabled.
struct TraceRef {
  unsigned long long* buf;
  unsigned int cap, tag;
};
TraceRef trace_ref(unsigned int tag);  // capi.cu: the buffer registered with hdpo_debug_set_trace, or {nullptr}
#ifndef HDPO_EMU
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_emit(const TraceRef& tr, unsigned long long t0, unsigned int block_linear,
                                           unsigned int aux = 0) {
  if (!tr.buf) return;
  const unsigned long long t1 = trace_now();
  unsigned int smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const unsigned int idx = atomicAdd(reinterpret_cast<unsigned int*>(tr.buf), 1u);
  if (idx < tr.cap) {
    unsigned long long* rec = tr.buf + 4 + 4ull * idx;
    rec[0] = t0;
    rec[1] = t1;
    rec[2] = static_cast<unsigned long long>(smid) | (static_cast<unsigned long long>(tr.tag) << 16);
    rec[3] = static_cast<unsigned long long>(block_linear) | (static_cast<unsigned long long>(aux) << 32);
  }
}
#else
inline unsigned long long trace_now() { return 0; }
inline void trace_emit(const TraceRef&, unsigned long long, unsigned int, unsigned int = 0) {}
#endif

constexpr int kWarp = 32;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace hdpo
