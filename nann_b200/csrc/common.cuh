// common.cuh -- error plumbing, launch accounting, host/device pointer staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/nann_b200.h"

namespace nann {

extern std::atomic<uint64_t> g_launches;
void set_error(const char* fmt, ...);
nann_status fail(nann_status code, const char* fmt, ...);

#define NANN_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::nann::fail(_e == cudaErrorMemoryAllocation ? NANN_RESOURCE_EXHAUSTED       \
                          : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) \
                              ? NANN_FAILED_PRECONDITION                                  \
                              : NANN_INTERNAL,                                            \
                          "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                      \
  } while (0)

#define NANN_TRY(expr)                    \
  do {                                    \
    nann_status _s = (expr);              \
    if (_s != NANN_OK) return _s;         \
  } while (0)

// every kernel launch in the library goes through this so nann_kernel_launch_count() is exact
#define NANN_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                      \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);             \
    ::nann::g_launches.fetch_add(1, std::memory_order_relaxed);             \
  } while (0)

nann_status require_device();  // NANN_FAILED_PRECONDITION when no CUDA device is usable
bool is_device_ptr(const void* p);

// A device view of caller memory: aliases device pointers, stages host pointers.
// Staging buffers come from cudaMalloc (freed in the destructor after a stream sync by the owner).
template <typename T>
struct DevIn {
  const T* d = nullptr;
  T* owned = nullptr;
  ~DevIn() { if (owned) cudaFree(owned); }
  nann_status init(const T* p, int64_t n, cudaStream_t st) {
    if (n <= 0 || p == nullptr) { d = nullptr; return NANN_OK; }
    if (is_device_ptr(p)) { d = p; return NANN_OK; }
    NANN_CUDA(cudaMalloc(&owned, (size_t)n * sizeof(T)));
    NANN_CUDA(cudaMemcpyAsync(owned, p, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, st));
    d = owned;
    return NANN_OK;
  }
};

// Output/in-out staging: device alias or device temp that is copied back by finish().
template <typename T>
struct DevOut {
  T* d = nullptr;
  T* owned = nullptr;
  T* host = nullptr;
  int64_t n = 0;
  ~DevOut() { if (owned) cudaFree(owned); }
  nann_status init(T* p, int64_t count, cudaStream_t st, bool copy_in) {
    n = count;
    if (count <= 0 || p == nullptr) { d = nullptr; return NANN_OK; }
    if (is_device_ptr(p)) { d = p; return NANN_OK; }
    host = p;
    NANN_CUDA(cudaMalloc(&owned, (size_t)count * sizeof(T)));
    if (copy_in) NANN_CUDA(cudaMemcpyAsync(owned, p, (size_t)count * sizeof(T), cudaMemcpyHostToDevice, st));
    d = owned;
    return NANN_OK;
  }
  // enqueue the copy back (caller syncs the stream)
  nann_status finish(cudaStream_t st, int64_t count = -1) {
    if (host && owned) {
      int64_t c = count < 0 ? n : count;
      if (c > 0) NANN_CUDA(cudaMemcpyAsync(host, owned, (size_t)c * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    return NANN_OK;
  }
  bool staged() const { return host != nullptr; }
};

template <typename T>
struct DevBuf {  // plain owned device allocation
  T* d = nullptr;
  int64_t n = 0;
  ~DevBuf() { if (d) cudaFree(d); }
  nann_status alloc(int64_t count) {
    if (d) { cudaFree(d); d = nullptr; }
    n = count;
    if (count <= 0) return NANN_OK;
    NANN_CUDA(cudaMalloc(&d, (size_t)count * sizeof(T)));
    return NANN_OK;
  }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nann
