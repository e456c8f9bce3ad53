// hdpo_internal.cuh - error plumbing shared by the host wrappers behind the C ABI.
#pragma once

#include <cstdarg>
#include <cstdio>

#include "hdpo_math.cuh"

namespace hdpo {

void set_error(const char* fmt, ...);  // capi.cu (thread-local message for hdpo_last_error)

#define HDPO_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::hdpo::set_error(__VA_ARGS__);      \
      return HDPO_E_INVALID;               \
    }                                      \
  } while (0)

#define HDPO_CUDA_OK(expr)                                                               \
  do {                                                                                   \
    cudaError_t hdpo_e_ = (expr);                                                        \
    if (hdpo_e_ != cudaSuccess) {                                                        \
      ::hdpo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(hdpo_e_), __FILE__, __LINE__); \
      return HDPO_E_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define HDPO_LAUNCH_OK()                                                                  \
  do {                                                                                    \
    cudaError_t hdpo_e_ = cudaGetLastError();                                             \
    if (hdpo_e_ != cudaSuccess) {                                                         \
      ::hdpo::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(hdpo_e_), __FILE__, __LINE__); \
      return HDPO_E_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

constexpr int kMaxNodes = 16;  // max warehouses / echelons handled by the per-scenario node loops

int validate_problem(const HdpoProblem* pb);

}  // namespace hdpo
