// capi.cu - library-level pieces of the C ABI (error string, launch counter, device query).
#include <atomic>
#include <cstring>

#include "hdpo_internal.cuh"

namespace hdpo {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static unsigned long long* g_trace_buf = nullptr;
static unsigned int g_trace_cap = 0;
TraceRef trace_ref(unsigned int tag) { return TraceRef{g_trace_buf, g_trace_cap, tag}; }

}  // namespace hdpo

// Debug hook: register (or clear with NULL) a device buffer of 4 + 4 * capacity uint64 that instrumented kernels
// append one {start ns, end ns, smid | tag << 16, block} record per CTA to; buf[0] (zeroed by the caller) counts.
extern "C" int hdpo_debug_set_trace(unsigned long long* buf, int64_t capacity) {
  hdpo::g_trace_buf = buf;
  hdpo::g_trace_cap = buf ? static_cast<unsigned int>(capacity) : 0u;
  return HDPO_OK;
}

extern "C" const char* hdpo_last_error(void) { return hdpo::g_err; }
extern "C" int hdpo_abi_version(void) { return HDPO_ABI_VERSION; }
extern "C" int64_t hdpo_kernel_launch_count(void) { return hdpo::g_launches.load(); }

extern "C" int hdpo_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char* name,
                                int32_t name_len) {
#ifdef HDPO_EMU
  if (sm_count) *sm_count = 1;
  if (cc_major) *cc_major = 0;
  if (cc_minor) *cc_minor = 0;
  if (name && name_len > 0) snprintf(name, name_len, "host-thread emulator (tests only)");
  return HDPO_OK;
#else
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    hdpo::set_error("no CUDA device: this library has no CPU fallback");
    return HDPO_E_NO_DEVICE;
  }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    hdpo::set_error("cudaGetDeviceProperties failed");
    return HDPO_E_NO_DEVICE;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (name && name_len > 0) snprintf(name, name_len, "%s", p.name);
  return HDPO_OK;
#endif
}
