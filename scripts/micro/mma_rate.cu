// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N and of where A
// comes from (shared memory descriptor vs TMEM), one CTA per SM, 512 back-to-back MMAs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int N, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) ((uint32_t*)sm)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(sm), b = smem_u32(sm + 16384);
    const uint32_t id = idesc(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t ks = (i & 3) * 32;
      const uint64_t bd = desc_sw128(b + ks);
      if (A_TMEM) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(tmem), "r"(tmem + 256 + (i & 3) * 8), "l"(bd), "r"(id), "r"(1u) : "memory");
      } else {
        const uint64_t ad = desc_sw128(a + ks);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(id), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool A_TMEM>
void run(const char* name, int grid) {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 512, smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(k<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, A_TMEM><<<grid, 128, smem>>>(d, iters);
  k<N, A_TMEM><<<grid, 128, smem>>>(d, iters);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s grid=%3d  issue %.1f cyc/mma   complete %.1f cyc/mma   (%s)\n", name, grid, (double)h[0] / iters, (double)h[1] / iters,
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<256, false>("SS  M128 N256 K16", grid);
    run<128, false>("SS  M128 N128 K16", grid);
    run<64, false>("SS  M128 N64  K16", grid);
    run<256, true>("TS  M128 N256 K16 (A in TMEM)", grid);
    run<64, true>("TS  M128 N64  K16 (A in TMEM)", grid);
  }
  return 0;
}
