// common.cuh -- error plumbing, launch accounting, host/device pointer staging.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
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

// every kernel launch in the library goes through this so nann_kernel_launch_count() is exact.  Launch-time
// failures (no sm_100a image for the device, a bad shared-memory or cluster configuration) are NOT sticky and a
// later cudaStreamSynchronize does not report them, so the launch result is checked right here.
#define NANN_LAUNCH_OR(on_error, kernel, grid, block, smem, stream, ...)                                  \
  do {                                                                                                    \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                           \
    ::nann::g_launches.fetch_add(1, std::memory_order_relaxed);                                           \
    cudaError_t _le = cudaPeekAtLastError();                                                              \
    if (_le != cudaSuccess) {                                                                             \
      cudaGetLastError();                                                                                 \
      return on_error(::nann::fail(_le == cudaErrorNoKernelImageForDevice ? NANN_FAILED_PRECONDITION      \
                                                                          : NANN_INTERNAL,                \
                                   "launch of %s failed: %s (%s:%d)", #kernel, cudaGetErrorString(_le),    \
                                   __FILE__, __LINE__));                                                  \
    }                                                                                                     \
  } while (0)
#define NANN_LAUNCH(kernel, grid, block, smem, stream, ...) \
  NANN_LAUNCH_OR(::nann::status_identity, kernel, grid, block, smem, stream, __VA_ARGS__)
static inline nann_status status_identity(nann_status s) { return s; }

nann_status require_device();  // NANN_FAILED_PRECONDITION when no CUDA device is usable
bool is_device_ptr(const void* p);

// Temporaries of the op-level entry points come from a small caching pool instead of cudaMalloc/cudaFree:
// both calls are device-wide synchronisation points and cost 0.1-5 ms each once NCCL has registered peer
// mappings (2 x B200: 4-9 ms per nann_merge_topk call before this pool).  A block becomes reusable only after a
// device synchronisation in release(); sizes are rounded up to a power of two >= 4 KB and at most 1 GiB per
// device stays cached.
struct DevPool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_blocks;          // (device, size class) -> block
  std::unordered_map<void*, std::pair<int, size_t>> live;            // block -> (device, size class)
  size_t cached_bytes = 0;
  static DevPool& get() { static DevPool* p = new DevPool(); return *p; }   // leaked on purpose: no teardown order issues
  static size_t size_class(size_t bytes) { size_t c = 4096; while (c < bytes) c <<= 1; return c; }
  cudaError_t alloc(void** out, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t c = size_class(bytes);
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.find({dev, c});
      if (it != free_blocks.end()) {
        *out = it->second;
        free_blocks.erase(it);
        cached_bytes -= c;
        live[*out] = {dev, c};
        return cudaSuccess;
      }
    }
    e = cudaMalloc(out, c);
    if (e == cudaErrorMemoryAllocation) {   // give the cache back and retry once
      cudaGetLastError();
      trim(0);
      e = cudaMalloc(out, c);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(mu); live[*out] = {dev, c}; }
    return e;
  }
  void release(void* p) {
    if (!p) return;
    // what cudaFree did implicitly: no kernel that may still use the block is in flight when it becomes reusable
    // (normal paths have already synchronised their stream, so this returns at once; error paths have not)
    cudaDeviceSynchronize();
    std::unique_lock<std::mutex> g(mu);
    auto it = live.find(p);
    if (it == live.end()) { g.unlock(); cudaFree(p); return; }
    const auto key = it->second;
    live.erase(it);
    if (cached_bytes + key.second > ((size_t)1 << 30)) { g.unlock(); cudaFree(p); return; }
    free_blocks.insert({key, p});
    cached_bytes += key.second;
  }
  void trim(size_t keep_bytes) {
    std::vector<void*> drop;
    {
      std::lock_guard<std::mutex> g(mu);
      while (cached_bytes > keep_bytes && !free_blocks.empty()) {
        auto it = free_blocks.begin();
        cached_bytes -= it->first.second;
        drop.push_back(it->second);
        free_blocks.erase(it);
      }
    }
    for (void* q : drop) cudaFree(q);
  }
};
template <typename T>
static inline cudaError_t pool_alloc(T** out, size_t bytes) { return DevPool::get().alloc((void**)out, bytes); }
static inline void pool_free(void* p) { DevPool::get().release(p); }

// A device view of caller memory: aliases device pointers, stages host pointers.
// Staging buffers come from the pool (handed back in the destructor after a stream sync by the owner).
template <typename T>
struct DevIn {
  const T* d = nullptr;
  T* owned = nullptr;
  ~DevIn() { if (owned) pool_free(owned); }
  nann_status init(const T* p, int64_t n, cudaStream_t st) {
    if (n <= 0 || p == nullptr) { d = nullptr; return NANN_OK; }
    if (is_device_ptr(p)) { d = p; return NANN_OK; }
    NANN_CUDA(pool_alloc(&owned, (size_t)n * sizeof(T)));
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
  ~DevOut() { if (owned) pool_free(owned); }
  nann_status init(T* p, int64_t count, cudaStream_t st, bool copy_in) {
    n = count;
    if (count <= 0 || p == nullptr) { d = nullptr; return NANN_OK; }
    if (is_device_ptr(p)) { d = p; return NANN_OK; }
    host = p;
    NANN_CUDA(pool_alloc(&owned, (size_t)count * sizeof(T)));
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
  ~DevBuf() { if (d) pool_free(d); }
  nann_status alloc(int64_t count) {
    if (d) { pool_free(d); d = nullptr; }
    n = count;
    if (count <= 0) return NANN_OK;
    NANN_CUDA(pool_alloc(&d, (size_t)count * sizeof(T)));
    return NANN_OK;
  }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nann
