// Micro-benchmark / bring-up: tcgen05.mma.cta_group::2 (M=256 over a CTA pair, B operand split over the two CTAs'
// shared memories) -- the mechanism that halves the shared-memory operand traffic per SM, which is what bounds the
// tensor-core scorer.  Checks one K=64 product exactly, then times back-to-back MMAs with and
// without a concurrent shared-memory write stream, against the cta_group::1 M128 N256 form.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_cta2 mma_cta2.cu && ./mma_cta2
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
  if (CG == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(id), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(id), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit_mc(uint32_t bar, uint16_t mask) {
  if (CG == 2)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// CG = 2: cluster of 2, leader issues M256 N256 K16 MMAs.  CG = 1: every CTA issues M128 N256 K16 on its own.
// stream != 0: warps 2,3 keep writing 16-B vectors into a scratch area of shared memory meanwhile.
template <int CG>
__global__ void __launch_bounds__(128, 1) k(long long* out, int* mismatches, int n_mma, int stream) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sAm = sm;                       // A: [128 rows][64 k] fp16 SW128 = 16 KB
  uint8_t* sBm = sm + 16384;               // B: CG=2 [128 n][64 k] (this CTA's half); CG=1 [256 n][64 k] = 32 KB
  uint8_t* scratch = sm + 65536;           // 64 KB write target of the traffic warps
  __shared__ uint64_t bar;
  __shared__ uint32_t slot_;
  __shared__ volatile int stop;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A[m][k] = (k == m % 64); B[n][k] = (n * 64 + k) % 251  => D[m][n] = B[n][m % 64]
  for (int i = tid; i < 128 * 64; i += 128) {
    const int r = i >> 6, kk = i & 63;
    const int m = (int)rank * 128 + r;
    *(__half*)(sAm + sw128_off(r, kk >> 3) + (kk & 7) * 2) = __float2half(kk == (m % 64) ? 1.f : 0.f);
  }
  const int n_rows_b = CG == 2 ? 128 : 256;
  for (int i = tid; i < n_rows_b * 64; i += 128) {
    const int r = i >> 6, kk = i & 63;
    const int n = (CG == 2 ? (int)rank * 128 : 0) + r;
    *(__half*)(sBm + sw128_off(r, kk >> 3) + (kk & 7) * 2) = __float2half((float)((n * 64 + kk) % 251));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot_)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot_)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot_;
  const uint32_t id = idesc(CG == 2 ? 256 : 128, 256);
  const bool issuer = (warp == 0 && lane == 0 && (CG == 1 || rank == 0));
  const uint16_t mask = CG == 2 ? 3 : 1;
  // ---- 1. one exact K=64 product
  if (issuer) {
    for (int ks = 0; ks < 4; ++ks) mma<CG>(tmem, desc_sw128(smem_u32(sAm) + ks * 32), desc_sw128(smem_u32(sBm) + ks * 32), id, ks ? 1u : 0u);
    commit_mc<CG>(smem_u32(&bar), mask);
  }
  wait_bar(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    int bad = 0;
    const int m = (int)rank * 128 + warp * 32 + lane;          // TMEM lane = row inside this CTA
    for (int c0 = 0; c0 < 256; c0 += 8) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) {
        const int n = c0 + j;
        const float want = (float)((n * 64 + (m % 64)) % 251);
        if (__uint_as_float(v[j]) != want) ++bad;
      }
    }
    if (bad) atomicAdd(mismatches, bad);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- 2. back-to-back MMAs
  if (warp >= 2 && stream) {                                   // shared-memory write traffic: 2 warps x 512 B per iteration
    uint4 z = make_uint4(tid, 1, 2, 3);
    int it = 0;
    while (!stop) {
      *(uint4*)(scratch + ((it * 1024 + (warp - 2) * 512 + lane * 16) & 65535)) = z;
      ++it;
    }
  }
  if (issuer) {
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) mma<CG>(tmem + (i & 1) * 256, desc_sw128(smem_u32(sAm) + (i & 3) * 32), desc_sw128(smem_u32(sBm) + (i & 3) * 32), id, 1u);
    commit_mc<CG>(smem_u32(&bar), mask);
    wait_bar(smem_u32(&bar), 1);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
    stop = 1;
  } else if (warp == 0 && lane == 0) {
    wait_bar(smem_u32(&bar), 1);                               // follower CTA (CG=2): the multicast commit lands here too
    stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else         asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int CG>
void run(int grid, int stream) {
  long long* d; int* mm;
  cudaMalloc(&d, 8); cudaMalloc(&mm, 4); cudaMemset(mm, 0, 4); cudaMemset(d, 0, 8);
  const int smem = 65536 + 65536 + 1024, n_mma = 400;
  cudaFuncSetAttribute(k<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k<CG>, d, mm, n_mma, stream);
  long long h = 0; int bad = -1;
  if (e == cudaSuccess) e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(&bad, mm, 4, cudaMemcpyDeviceToHost);
  printf("cta_group::%d grid=%3d smem-traffic=%d : mismatches=%d  %6.1f cyc/MMA (M%d N256 K16)  %s\n", CG, grid, stream, bad,
         (double)h / n_mma, CG == 2 ? 256 : 128, cudaGetErrorString(e));
  cudaFree(d); cudaFree(mm);
}

int main() {
  for (int stream : {0, 1}) {
    run<1>(1, stream); run<1>(148, stream);
    run<2>(2, stream); run<2>(148, stream);
  }
  return 0;
}
