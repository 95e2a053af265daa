// xc_common.cuh — shared host-side helpers of libxcape_b200 (error plumbing, launch counter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/xcape_b200.h"

namespace xc {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define XC_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      return ::xc::fail(XCAPE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    }                                                                                      \
  } while (0)

#define XC_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    ::xc::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      return ::xc::fail(XCAPE_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    }                                                                                      \
  } while (0)

// event pair around the dominant kernel of a call (xcape_cuda_time_kernels / xcape_cuda_last_kernel_ms); no-ops when off
void kernel_timer_begin(cudaStream_t s);
void kernel_timer_end(cudaStream_t s);

inline size_t esize(int dtype) { return dtype == XCAPE_F64 ? 8 : 4; }

// entry points of the per-TU kernel launchers
int launch_cape_faithful(const struct CapeArgs& a, int source, int adiabat, bool p1d, cudaStream_t s);
size_t cape_sort_scratch_bytes(int64_t ncol, int nlev);
int launch_cape_fast(const struct CapeArgs& a, int source, int adiabat, bool p1d, cudaStream_t s);
int launch_exner_table(const float* p_hpa, float* pi, int nlev, cudaStream_t s);
int launch_cape_fast_relaxed(const struct CapeArgs& a, int source, int adiabat, bool p1d, cudaStream_t s);

}  // namespace xc
