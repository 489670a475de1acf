// Micro-benchmark: does tcgen05.mma throughput depend on HOW the MMAs are issued and on concurrent TMA traffic?
//   ISSUE 0: single thread (threadIdx.x == 0) loop            -> ptxas wraps each UTCHMMA in a waterfall loop
//   ISSUE 1: warp-uniform loop, elect.sync leader, unrolled x8 -> back-to-back UTCHMMA with uniform registers
//   TMA   1: a second warp streams 32 KB cp.async.bulk copies global -> shared (4 in flight) meanwhile
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_burst mma_burst.cu && ./mma_burst
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t id) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(id), "r"(1u) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

#define POLL_ALL_LANES(P) ((P) > 0)
constexpr int A_BYTES = 16384, B_BYTES = 32768, STAGE = 32768, NSLOT = 4;

template <int N, int ISSUE, int TMA, int DSPLIT, int RND = 0, int POLL = 0>
__global__ void __launch_bounds__(128 + 32 * POLL, 1) k(long long* out, int iters, const uint8_t* src, int copies) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, tbar[NSLOT], never;
  __shared__ uint32_t slot;
  __shared__ volatile int mma_done;
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += 128) {
    uint32_t v = 0x3c003c00u;  // fp16 1.0
    if (RND) {   // random fp16 in roughly [-2, 2): random sign + mantissa, exponent 0x3c..0x3f
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x83ff83ffu) | 0x3c003c00u | ((h >> 4) & 0x04000400u);
    }
    ((uint32_t*)sm)[i] = v;
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)));
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mma_done = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a = smem_u32(sm), b = smem_u32(sm + A_BYTES);
  const uint32_t id = idesc(128, N);
  if (warp == 0) {
    long long t0 = clock64(), t1;
    if (ISSUE == 0) {
      if (lane == 0) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t ks = (i & 3) * 32;
          mma(tmem + (DSPLIT ? (i & 1) * 256 : 0), desc_sw128(a + ks), desc_sw128(b + ks), id);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
    } else {
      for (int i = 0; i < iters; i += 8) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t ks = (j & 3) * 32;
            mma(tmem + (DSPLIT ? (j & 1) * 256 : 0), desc_sw128(a + ks), desc_sw128(b + ks), id);
          }
        }
        __syncwarp();
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    t1 = clock64();
    if (lane == 0) {
      wait_bar(smem_u32(&bar), 0);
      long long t2 = clock64();
      mma_done = 1;
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (warp >= 4) {   // pollers: spin on a barrier that never completes (POLL: 1 = all lanes, 2 = lane 0 only)
    if (POLL_ALL_LANES(POLL) || lane == 0) {
      while (!mma_done) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&never)), "r"(0u) : "memory");
      }
    }
  } else if (warp == 1 && TMA) {
    if (lane == 0) {
      uint8_t* ring = sm + A_BYTES + B_BYTES;
      long long t0 = clock64();
      int done_copies = 0;
      for (int i = 0; i < copies && !mma_done; ++i) {
        const int s = i % NSLOT;
        if (i >= NSLOT) wait_bar(smem_u32(&tbar[s]), ((i / NSLOT) - 1) & 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[s])), "r"((uint32_t)STAGE) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(ring + s * STAGE)), "l"(src + (size_t)((i * 7 + blockIdx.x) % 64) * STAGE), "r"((uint32_t)STAGE),
                       "r"(smem_u32(&tbar[s])) : "memory");
        done_copies = i + 1;
      }
      // drain
      for (int i = max(0, done_copies - NSLOT); i < done_copies; ++i) wait_bar(smem_u32(&tbar[i % NSLOT]), (i / NSLOT) & 1);
      long long t1 = clock64();
      if (blockIdx.x == 0) { out[2] = t1 - t0; out[3] = done_copies; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int ISSUE, int TMA, int DSPLIT, int RND = 0, int POLL = 0>
void run(const char* name, int grid, const uint8_t* src) {
  long long* d;
  cudaMalloc(&d, 32);
  cudaMemset(d, 0, 32);
  const int iters = 512, smem = A_BYTES + B_BYTES + NSLOT * STAGE + 1024;
  cudaFuncSetAttribute(k<N, ISSUE, TMA, DSPLIT, RND, POLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, ISSUE, TMA, DSPLIT, RND, POLL><<<grid, 128 + 32 * POLL, smem>>>(d, iters, src, 100000);
  k<N, ISSUE, TMA, DSPLIT, RND, POLL><<<grid, 128 + 32 * POLL, smem>>>(d, iters, src, 100000);
  long long h[4] = {0, 0, 0, 0};
  cudaError_t e = cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("%-44s grid=%3d issue %6.1f  complete %6.1f cyc/mma", name, grid, (double)h[0] / iters, (double)h[1] / iters);
  if (TMA) printf("   tma %5.1f B/cyc (%lld copies)", h[2] ? (double)h[3] * STAGE / (double)h[2] : 0.0, h[3]);
  printf("  (%s)\n", cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  uint8_t* src;
  cudaMalloc(&src, 64 * STAGE);
  cudaMemset(src, 0, 64 * STAGE);
  for (int grid : {1, 148}) {
    run<256, 0, 0, 0>("N256 single-thread issue", grid, src);
    run<256, 1, 0, 0>("N256 elect burst x8", grid, src);
    run<256, 1, 0, 1>("N256 elect burst x8, alternating D", grid, src);
    run<256, 0, 1, 0>("N256 single-thread issue + TMA stream", grid, src);
    run<256, 1, 1, 0>("N256 elect burst x8 + TMA stream", grid, src);
    run<256, 0, 0, 0, 1>("N256 single-thread issue, RANDOM data", grid, src);
    run<256, 1, 0, 0, 1>("N256 elect burst x8, RANDOM data", grid, src);
    run<256, 1, 1, 0, 1>("N256 elect burst x8 + TMA, RANDOM data", grid, src);
    run<128, 1, 0, 0, 1>("N128 elect burst x8, RANDOM data", grid, src);
    run<256, 1, 1, 0, 1, 8>("N256 burst + TMA + 8 polling warps", grid, src);
    run<256, 0, 1, 0, 1, 8>("N256 single-thread + TMA + 8 polling warps", grid, src);
    run<256, 1, 0, 0, 1, 8>("N256 burst + 8 polling warps", grid, src);
    run<256, 1, 0, 0, 1, 2>("N256 burst + 2 polling warps", grid, src);
    run<128, 0, 0, 0>("N128 single-thread issue", grid, src);
    run<128, 1, 0, 0>("N128 elect burst x8", grid, src);
    run<128, 1, 1, 0>("N128 elect burst x8 + TMA stream", grid, src);
  }
  return 0;
}
