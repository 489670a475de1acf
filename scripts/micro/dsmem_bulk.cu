// Micro-benchmark: DSMEM bulk copy (cp.async.bulk.shared::cluster.shared::cta) between the two CTAs of a cluster,
// both directions at once (the scorer v8 h1 exchange).  Reports bytes/clk per direction for a few copy sizes and
// numbers of copies in flight.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bulk dsmem_bulk.cu && ./dsmem_bulk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// each CTA sends `n` chunks of `bytes` to the peer; `depth` chunks in flight (receiver-side barriers, one per in-flight slot)
__global__ void __launch_bounds__(128, 1) k(long long* out, int n, int bytes, int depth, int bidir) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], ack[8];
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1;
  uint8_t* src = sm;                 // 64 KB
  uint8_t* dst = sm + 65536;         // depth x bytes (<= 128 KB)
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ack[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 16384; i += 128) ((uint32_t*)src)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  const bool sender = bidir || rank == 0;
  const bool receiver = bidir || rank == 1;
  long long t0 = clock64();
  if (threadIdx.x == 0 && sender) {
    // sender: chunk i goes to the peer's dst slot i % depth once the peer acked chunk i - depth
    const uint32_t rdst = mapa_u32(smem_u32(dst), peer), rfull = mapa_u32(smem_u32(&full[0]), peer);
    for (int i = 0; i < n; ++i) {
      const int s = i % depth;
      if (i >= depth) wait_bar(smem_u32(&ack[s]), ((i / depth) - 1) & 1);
      asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(rfull + 8 * s), "r"((uint32_t)bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(rdst + s * bytes), "r"(smem_u32(src) + (i & 1) * bytes % 65536), "r"((uint32_t)bytes), "r"(rfull + 8 * s) : "memory");
    }
  }
  if (threadIdx.x == 32 && receiver) {
    const uint32_t rack = mapa_u32(smem_u32(&ack[0]), peer);
    for (int i = 0; i < n; ++i) {
      const int s = i % depth;
      wait_bar(smem_u32(&full[s]), (i / depth) & 1);
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rack + 8 * s) : "memory");
    }
    if (blockIdx.x == (bidir ? 0 : 1)) out[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 65536 + 131072 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {2, 148})
    for (int bidir : {0, 1})
      for (int bytes : {32768, 8192, 2048})
        for (int depth : {1, 2, 4}) {
          if ((long long)bytes * depth > 131072) continue;
          const int n = 4 * 1024 * 1024 / bytes;
          cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, k, d, n, bytes, depth, bidir);
          cudaLaunchKernelEx(&cfg, k, d, n, bytes, depth, bidir);
          long long h = 0; cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("grid=%3d bidir=%d chunk=%6d B depth=%d : %8.1f cyc/chunk  %6.2f B/clk per direction (%s)\n", grid, bidir, bytes, depth,
                 (double)h / n, (double)bytes * n / (double)h, cudaGetErrorString(e));
        }
  return 0;
}
